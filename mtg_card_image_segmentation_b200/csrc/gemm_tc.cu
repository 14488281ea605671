// Pointwise (1x1) and 3x3 convolutions as implicit GEMMs on the sm_100a tensor cores.
//
//   out[M][N] = act( (A[M][K] . W[N][K]^T) * scale[N] + shift[N] ) (+ residual[M][N])
//
// * A is the NHWC bf16 activation tensor itself (rows = pixels, K = channels): no im2col buffer exists.
//   1x1: a 2-D TMA tile [128 rows][64 ch].  3x3: nine shifted 4-D TMA boxes {64 ch, W, HB rows, NB images}
//   per 64-channel slab; the TMA out-of-bounds zero fill IS the convolution padding.
// * W is [N][taps*K] bf16 (K contiguous), one 2-D TMA tile [BN rows][64] per stage.
// * tcgen05.mma (kind::f16, bf16 x bf16 -> fp32) with both operands K-major in 128B-swizzled shared
//   memory; accumulators live in TMEM, double buffered so the epilogue of tile i overlaps the loads and
//   MMAs of tile i+1.  Persistent CTAs walk the tile list round-robin.
// * warp roles: w0 = TMA producer, w1 = MMA issuer (+ TMEM owner), w2..5 = epilogue (one TMEM lane
//   quarter each), w6..9 (only with kAScale) = squeeze-excite prologue that rescales the A tile in
//   shared memory per (image, channel) before the MMA reads it.
//
// Replaces, for this path, what the reference delegates to cuDNN/oneDNN: nn.Conv2d 1x1 in
// tv:models/mobilenetv3.py:71-80,101-105,179-187 and the head's 3x3 (train/model.py:110), with the
// eval-mode BatchNorm + activation (+ residual, tv:models/mobilenetv3.py:111-115) fused in the epilogue.
#include <cuda.h>

#include <mutex>

#include "ops.h"
#include "ptx.cuh"

namespace mtgseg {

namespace {

constexpr int BM = 128;           // UMMA M
constexpr int BK = 64;            // 64 bf16 = one 128-byte swizzle row
constexpr int A_STAGE_BYTES = BM * BK * 2;
constexpr int MAX_STAGES = 6;

struct GemmKParams {
  int M, N, K;
  int BN, n_tiles, m_tiles;
  int num_kb, kb_per_tap, ksteps_last;
  int stages, tmem_cols, a_bytes;  // a_bytes: bytes one A TMA box delivers
  const float* scale;
  const float* shift;
  int act;
  const bf16* residual;
  bf16* out;
  const float* a_scale;
  int hw;
  // 3x3 geometry
  int B, H, W, HB, NB, h_tiles;
};

template <bool kConv3x3, bool kAScale>
__global__ void __launch_bounds__(kAScale ? 320 : 192, 1)
conv_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmKParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int S = p.stages;
  const int b_stage_bytes = p.BN * BK * 2;
  uint8_t* sA = smem;
  uint8_t* sB = smem + S * A_STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB + S * b_stage_bytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + MAX_STAGES;
  uint64_t* xform = bars + 2 * MAX_STAGES;
  uint64_t* tfull = bars + 3 * MAX_STAGES;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) {
      ptx::mbar_init(&full[s], 1);
      ptx::mbar_init(&empty[s], 1);
      ptx::mbar_init(&xform[s], 128);
    }
    for (int b = 0; b < 2; ++b) {
      ptx::mbar_init(&tfull[b], 1);
      ptx::mbar_init(&tempty[b], 128);
    }
    ptx::fence_mbar_init();
  }
  if (warp == 0 && lane == 0) {
    ptx::tma_prefetch_desc(&tmA);
    ptx::tma_prefetch_desc(&tmB);
  }
  if (warp == 1) {
    ptx::tmem_alloc(tmem_slot, p.tmem_cols);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t buf_stride = p.tmem_cols >> 1;
  const int total_tiles = p.m_tiles * p.n_tiles;

  if (warp == 0) {
    // =============================== TMA producer ===============================
    if (lane == 0) {
      uint32_t it = 0;
      const uint32_t stage_bytes = p.a_bytes + b_stage_bytes;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int m_tile = tile / p.n_tiles, n_tile = tile - m_tile * p.n_tiles;
        int n0 = 0, y0 = 0;
        if (kConv3x3) {
          n0 = (m_tile / p.h_tiles) * p.NB;
          y0 = (m_tile % p.h_tiles) * p.HB;
        }
        for (int kb = 0; kb < p.num_kb; ++kb, ++it) {
          const int s = it % S;
          const uint32_t ph = (it / S) & 1;
          ptx::mbar_wait(&empty[s], ph ^ 1);
          ptx::mbar_arrive_expect_tx(&full[s], stage_bytes);
          if (kConv3x3) {
            const int tap = kb / p.kb_per_tap, kc = kb - tap * p.kb_per_tap;
            ptx::tma_load_4d(sA + s * A_STAGE_BYTES, &tmA, &full[s], kc * BK, tap % 3 - 1, y0 + tap / 3 - 1, n0);
            ptx::tma_load_2d(sB + s * b_stage_bytes, &tmB, &full[s], tap * p.K + kc * BK, n_tile * p.BN);
          } else {
            ptx::tma_load_2d(sA + s * A_STAGE_BYTES, &tmA, &full[s], kb * BK, m_tile * BM);
            ptx::tma_load_2d(sB + s * b_stage_bytes, &tmB, &full[s], kb * BK, n_tile * p.BN);
          }
        }
      }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer ===============================
    if (lane == 0) {
      const uint32_t idesc = ptx::umma_idesc_bf16(BM, p.BN);
      uint32_t it = 0, tc = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++tc) {
        const uint32_t buf = tc & 1, aph = (tc >> 1) & 1;
        ptx::mbar_wait(&tempty[buf], aph ^ 1);
        ptx::tc_fence_after();
        const uint32_t d = tmem_base + buf * buf_stride;
        for (int kb = 0; kb < p.num_kb; ++kb, ++it) {
          const int s = it % S;
          const uint32_t ph = (it / S) & 1;
          ptx::mbar_wait(kAScale ? &xform[s] : &full[s], ph);
          ptx::tc_fence_after();
          const int kc = kb % p.kb_per_tap;
          const int ksteps = (kc == p.kb_per_tap - 1) ? p.ksteps_last : (BK / 16);
          const uint64_t adesc = ptx::umma_desc_sw128_kmajor(ptx::smem_u32(sA + s * A_STAGE_BYTES));
          const uint64_t bdesc = ptx::umma_desc_sw128_kmajor(ptx::smem_u32(sB + s * b_stage_bytes));
          for (int k = 0; k < ksteps; ++k)  // +32 B along K inside the swizzle row == +2 in the address field
            ptx::umma_bf16(d, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
          ptx::umma_commit(&empty[s]);  // frees the smem stage once these MMAs have read it
        }
        ptx::umma_commit(&tfull[buf]);  // accumulator complete -> epilogue
      }
    }
  } else if (warp < 6) {
    // =============================== epilogue ===============================
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    const int r = q * 32 + lane;
    uint32_t tc = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++tc) {
      const int m_tile = tile / p.n_tiles, n_tile = tile - m_tile * p.n_tiles;
      const uint32_t buf = tc & 1, aph = (tc >> 1) & 1;
      long long m;
      bool valid;
      if (kConv3x3) {
        const int n0 = (m_tile / p.h_tiles) * p.NB, y0 = (m_tile % p.h_tiles) * p.HB;
        const int wx = r % p.W, t = r / p.W, hb = t % p.HB, nb = t / p.HB;
        valid = (nb < p.NB) && (y0 + hb < p.H) && (n0 + nb < p.B);
        m = (static_cast<long long>(n0 + nb) * p.H + (y0 + hb)) * p.W + wx;
      } else {
        m = static_cast<long long>(m_tile) * BM + r;
        valid = m < p.M;
      }
      ptx::mbar_wait(&tfull[buf], aph);
      ptx::tc_fence_after();
      const uint32_t taddr = tmem_base + buf * buf_stride + (static_cast<uint32_t>(q * 32) << 16);
      bf16* orow = p.out + m * p.N;
      const bf16* rrow = p.residual ? p.residual + m * p.N : nullptr;
      for (int c0 = 0; c0 < p.BN; c0 += 16) {
        uint32_t v[16];
        ptx::tmem_ld16(taddr + c0, v);
        ptx::tmem_ld_wait();
        const int nbase = n_tile * p.BN + c0;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int n = nbase + h * 8;
          if (valid && n < p.N) {  // N % 8 == 0, so groups of 8 never straddle N
            float f[8];
            const float4 sc0 = p.scale ? __ldg(reinterpret_cast<const float4*>(p.scale + n)) : make_float4(1, 1, 1, 1);
            const float4 sc1 = p.scale ? __ldg(reinterpret_cast<const float4*>(p.scale + n + 4)) : make_float4(1, 1, 1, 1);
            const float4 sh0 = p.shift ? __ldg(reinterpret_cast<const float4*>(p.shift + n)) : make_float4(0, 0, 0, 0);
            const float4 sh1 = p.shift ? __ldg(reinterpret_cast<const float4*>(p.shift + n + 4)) : make_float4(0, 0, 0, 0);
            const float sc[8] = {sc0.x, sc0.y, sc0.z, sc0.w, sc1.x, sc1.y, sc1.z, sc1.w};
            const float sh[8] = {sh0.x, sh0.y, sh0.z, sh0.w, sh1.x, sh1.y, sh1.z, sh1.w};
#pragma unroll
            for (int j = 0; j < 8; ++j) f[j] = apply_act(fmaf(__uint_as_float(v[h * 8 + j]), sc[j], sh[j]), p.act);
            if (rrow) {
              float rf[8];
              unpack8(ldg16(rrow + n), rf);
#pragma unroll
              for (int j = 0; j < 8; ++j) f[j] += rf[j];
            }
            *reinterpret_cast<uint4*>(orow + n) = pack8(f);
          }
        }
      }
      ptx::tc_fence_before();
      ptx::mbar_arrive(&tempty[buf]);
    }
  } else if (kAScale) {
    // =============================== squeeze-excite prologue on the A tile ===============================
    const int t = threadIdx.x - 192;
    const int chunk = t & 7;  // physical 16-byte chunk inside the 128-byte row
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int m_tile = tile / p.n_tiles;
      for (int kb = 0; kb < p.num_kb; ++kb, ++it) {
        const int s = it % S;
        const uint32_t ph = (it / S) & 1;
        ptx::mbar_wait(&full[s], ph);
        uint8_t* base = sA + s * A_STAGE_BYTES;
#pragma unroll 4
        for (int j = 0; j < 8; ++j) {
          const int row = (t >> 3) + 16 * j;
          const long long m = static_cast<long long>(m_tile) * BM + row;
          const int k = kb * BK + ((chunk ^ (row & 7)) << 3);  // undo the 128B swizzle: logical channel of this chunk
          if (m < p.M && k < p.K) {
            const int img = static_cast<int>(m / p.hw);
            const float* sp = p.a_scale + static_cast<size_t>(img) * p.K + k;
            const float4 s0 = __ldg(reinterpret_cast<const float4*>(sp));
            const float4 s1 = __ldg(reinterpret_cast<const float4*>(sp + 4));
            uint4* ptr = reinterpret_cast<uint4*>(base + row * 128 + chunk * 16);
            float f[8];
            unpack8(*ptr, f);
            f[0] *= s0.x; f[1] *= s0.y; f[2] *= s0.z; f[3] *= s0.w;
            f[4] *= s1.x; f[5] *= s1.y; f[6] *= s1.z; f[7] *= s1.w;
            *ptr = pack8(f);
          }
        }
        ptx::fence_proxy_async_smem();
        ptx::mbar_arrive(&xform[s]);
      }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

// bf16 tensor map, 128-byte swizzle, zero OOB fill.  dims/strides innermost first; strides in bytes for dims 1..
int make_map(CUtensorMap* map, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
             const uint32_t* box) {
  EncodeTiledFn fn = get_encode_fn();
  MTG_REQUIRE(fn != nullptr, MTG_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t gdim[5];
  cuuint64_t gstr[4];
  cuuint32_t bdim[5], estr[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bdim[i] = box[i];
    estr[i] = 1;
    if (i > 0) gstr[i - 1] = strides_bytes[i - 1];
  }
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, const_cast<void*>(base), gdim, gstr, bdim, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  MTG_REQUIRE(r == CUDA_SUCCESS, MTG_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d (rank %d, dim0 %llu)",
              static_cast<int>(r), rank, static_cast<unsigned long long>(dims[0]));
  return MTG_OK;
}

int g_num_sms = 0;
int num_sms() {
  if (g_num_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_num_sms <= 0) g_num_sms = 148;
  }
  return g_num_sms;
}

template <bool C3, bool AS>
int launch_variant(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmKParams& kp, int grid, size_t smem,
                   cudaStream_t st) {
  static bool configured = false;  // per template instantiation
  if (!configured) {
    MTG_CUDA(cudaFuncSetAttribute(conv_gemm_kernel<C3, AS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    configured = true;
  }
  conv_gemm_kernel<C3, AS><<<grid, AS ? 320 : 192, smem, st>>>(tmA, tmB, kp);
  MTG_LAUNCH_CHECK();
  return MTG_OK;
}

}  // namespace

int launch_conv_gemm(const ConvGemmArgs& g, cudaStream_t st) {
  MTG_REQUIRE(g.a && g.w && g.out, MTG_ERR_ARG, "conv_gemm: null pointer");
  MTG_REQUIRE(g.M > 0 && g.N > 0 && g.K > 0, MTG_ERR_ARG, "conv_gemm: bad shape M=%d N=%d K=%d", g.M, g.N, g.K);
  MTG_REQUIRE(g.K % 8 == 0 && g.N % 8 == 0, MTG_ERR_UNSUPPORTED,
              "conv_gemm: channel counts must be multiples of 8 (K=%d N=%d)", g.K, g.N);
  MTG_REQUIRE(!(g.conv3x3 && g.a_scale), MTG_ERR_UNSUPPORTED, "conv_gemm: a_scale with conv3x3 is not supported");
  MTG_REQUIRE(!g.a_scale || g.hw > 0, MTG_ERR_ARG, "conv_gemm: a_scale needs hw");

  GemmKParams kp{};
  kp.M = g.M; kp.N = g.N; kp.K = g.K;
  kp.n_tiles = ceil_div(g.N, 256);
  kp.BN = static_cast<int>(align_up(ceil_div(g.N, kp.n_tiles), 16));
  kp.kb_per_tap = ceil_div(g.K, BK);
  kp.num_kb = kp.kb_per_tap * (g.conv3x3 ? 9 : 1);
  kp.ksteps_last = ceil_div(g.K - (kp.kb_per_tap - 1) * BK, 16);
  kp.scale = g.scale; kp.shift = g.shift; kp.act = g.act;
  kp.residual = g.residual; kp.out = g.out; kp.a_scale = g.a_scale; kp.hw = g.hw;
  int tmem = 32;
  while (tmem < 2 * kp.BN) tmem <<= 1;
  kp.tmem_cols = tmem;

  CUtensorMap tmA, tmB;
  if (g.conv3x3) {
    MTG_REQUIRE(g.B > 0 && g.H > 0 && g.W > 0 && static_cast<long long>(g.B) * g.H * g.W == g.M, MTG_ERR_ARG,
                "conv_gemm 3x3: geometry mismatch");
    MTG_REQUIRE(g.W <= BM, MTG_ERR_UNSUPPORTED, "conv_gemm 3x3: feature-map width %d > %d unsupported", g.W, BM);
    // pick the TMA box {W, HB rows, NB images} with the fewest tiles (== best use of the 128 MMA rows)
    long long best_tiles = -1;
    for (int hb = 1; hb <= g.H && hb * g.W <= BM; ++hb) {
      int nb = BM / (hb * g.W);
      if (nb > g.B) nb = g.B;
      if (nb > 256) nb = 256;
      const long long tiles = static_cast<long long>(ceil_div(g.H, hb)) * ceil_div(g.B, nb);
      if (best_tiles < 0 || tiles < best_tiles) { best_tiles = tiles; kp.HB = hb; kp.NB = nb; }
    }
    kp.B = g.B; kp.H = g.H; kp.W = g.W;
    kp.h_tiles = ceil_div(g.H, kp.HB);
    kp.m_tiles = kp.h_tiles * ceil_div(g.B, kp.NB);
    kp.a_bytes = kp.NB * kp.HB * g.W * BK * 2;
    const uint64_t dims[4] = {(uint64_t)g.K, (uint64_t)g.W, (uint64_t)g.H, (uint64_t)g.B};
    const uint64_t strides[3] = {(uint64_t)g.K * 2, (uint64_t)g.W * g.K * 2, (uint64_t)g.H * g.W * g.K * 2};
    const uint32_t box[4] = {BK, (uint32_t)g.W, (uint32_t)kp.HB, (uint32_t)kp.NB};
    int rc = make_map(&tmA, g.a, 4, dims, strides, box);
    if (rc) return rc;
    const uint64_t wd[2] = {(uint64_t)g.K * 9, (uint64_t)g.N};
    const uint64_t ws[1] = {(uint64_t)g.K * 9 * 2};
    const uint32_t wb[2] = {BK, (uint32_t)kp.BN};
    rc = make_map(&tmB, g.w, 2, wd, ws, wb);
    if (rc) return rc;
  } else {
    kp.m_tiles = ceil_div(g.M, BM);
    kp.a_bytes = A_STAGE_BYTES;
    const uint64_t dims[2] = {(uint64_t)g.K, (uint64_t)g.M};
    const uint64_t strides[1] = {(uint64_t)g.K * 2};
    const uint32_t box[2] = {BK, BM};
    int rc = make_map(&tmA, g.a, 2, dims, strides, box);
    if (rc) return rc;
    const uint64_t wd[2] = {(uint64_t)g.K, (uint64_t)g.N};
    const uint64_t ws[1] = {(uint64_t)g.K * 2};
    const uint32_t wb[2] = {BK, (uint32_t)kp.BN};
    rc = make_map(&tmB, g.w, 2, wd, ws, wb);
    if (rc) return rc;
  }

  const int stage_bytes = A_STAGE_BYTES + kp.BN * BK * 2;
  int stages = 4;
  if (kp.num_kb >= 8) stages = 6;
  while (stages > 2 && stages * stage_bytes > 200 * 1024) --stages;
  kp.stages = stages;
  const size_t need = static_cast<size_t>(stages) * stage_bytes + 1024 /*align*/ + 256 /*barriers*/;
  MTG_REQUIRE(need <= 227 * 1024, MTG_ERR_UNSUPPORTED, "conv_gemm: tile needs %zu B shared memory", need);
  int per_sm = static_cast<int>((227 * 1024) / need);
  if (per_sm > 512 / kp.tmem_cols) per_sm = 512 / kp.tmem_cols;
  if (per_sm > 4) per_sm = 4;
  if (per_sm < 1) per_sm = 1;
  // pad the request so that never more than per_sm CTAs share an SM (their TMEM allocations always fit)
  size_t smem = need;
  const size_t floor_for_cap = (227 * 1024) / (per_sm + 1) + 1;
  if (smem < floor_for_cap) smem = floor_for_cap;
  const long long total_tiles = static_cast<long long>(kp.m_tiles) * kp.n_tiles;
  int grid = num_sms() * per_sm;
  if (grid > total_tiles) grid = static_cast<int>(total_tiles);

  if (g.conv3x3) return launch_variant<true, false>(tmA, tmB, kp, grid, smem, st);
  if (g.a_scale) return launch_variant<false, true>(tmA, tmB, kp, grid, smem, st);
  return launch_variant<false, false>(tmA, tmB, kp, grid, smem, st);
}

}  // namespace mtgseg
