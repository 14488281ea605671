// Weight gradient of the 1x1 / 3x3 convolutions on the sm_100a tensor cores.
//
//   dW[n][k][tap] += sum over pixels m of  dz[m][n] * x[shift_tap(m)][k]          (fp32, reference OIHW layout)
//
// As a GEMM the reduction runs over PIXELS, so both operands are consumed "MN-major": the NHWC activations are
// already [pixel][channel] with channels contiguous, i.e. exactly the transposed operand tcgen05 can read through an
// MN-major shared-memory descriptor -- no transpose pass, no im2col.  One TMA box = 64 pixels x 64 channels (128-byte
// swizzled rows); A' (dz, 128 output channels) is two boxes, B' (x, up to 256 input channels) up to four; the 3x3
// case shifts the x box by (dx, dy) and lets the TMA zero-fill the padding.
// Split-K: a CTA owns one output tile and a run of 64-pixel segments inside ONE image, accumulates in TMEM and adds
// its partial tile to dW with fp32 atomics (red.global); the squeeze-excite gate s[img][k] the forward applied to x
// is applied to the finished accumulator instead (it is constant inside an image).
// Replaces the CUDA-core wgrad of train_conv.cu for channel counts >= 64.
#include <cuda.h>

#include "ops.h"
#include "ptx.cuh"

namespace mtgseg {

int make_tma_map_bf16(CUtensorMap* map, const void* base, int rank, const unsigned long long* dims,
                      const unsigned long long* strides_bytes, const unsigned* box, int kbox);  // gemm_tc.cu

namespace {

constexpr int SEG = 64;                 // pixels per pipeline stage
constexpr int BOX_BYTES = SEG * 128;    // one [64 px][64 ch] box
constexpr int STAGES = 4;

struct WgTcP {
  int N, K, taps, BNp, nboxes_b, mt_tiles, nt_tiles;
  int segs_per_img, segs_per_cta, runs_per_img, rows_per_seg;  // rows_per_seg: valid smem rows per box (64, or HB*W for 3x3)
  int imgs_per_cta, B;  // without a squeeze-excite gate a CTA may keep accumulating over several images
  int H, W, HB;
  float* dw;
  const float* a_scale;
  uint32_t tmem_cols;
};

template <bool kConv3x3>
__global__ void __launch_bounds__(192, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmDz, const __grid_constant__ CUtensorMap tmX, const WgTcP p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int a_bytes = 2 * BOX_BYTES, b_bytes = p.nboxes_b * BOX_BYTES, stage_bytes = a_bytes + b_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * stage_bytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + STAGES;
  uint64_t* done = bars + 2 * STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // decode the work item
  int tile = blockIdx.x;
  const int tap = tile % p.taps; tile /= p.taps;
  const int nt = tile % p.nt_tiles, mt = tile / p.nt_tiles;
  const int img0 = (blockIdx.y / p.runs_per_img) * p.imgs_per_cta, run = blockIdx.y % p.runs_per_img;
  const int nimg = min(p.imgs_per_cta, p.B - img0);
  const int seg0 = run * p.segs_per_cta, seg1 = min(p.segs_per_img, seg0 + p.segs_per_cta);
  const int n0 = mt * 128, k0 = nt * p.BNp;
  const int segs = seg1 - seg0;
  const int nseg = segs * nimg;  // pipeline iterations: (image, segment) pairs
  const int img = img0;

  if (kConv3x3) {  // boxes deliver fewer than 64 rows: the unused rows take part in the reduction and must be zero
    for (int i = threadIdx.x; i < STAGES * stage_bytes / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
    ptx::fence_proxy_async_smem();
  }
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { ptx::mbar_init(&full[s], 1); ptx::mbar_init(&empty[s], 1); }
    ptx::mbar_init(done, 1);
    ptx::fence_mbar_init();
  }
  if (warp == 1) { ptx::tmem_alloc(tmem_slot, p.tmem_cols); ptx::tmem_relinquish(); }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      const uint32_t tx = static_cast<uint32_t>((2 + p.nboxes_b) * p.rows_per_seg * 128);
      for (int i = 0; i < nseg; ++i) {
        const int s = i % STAGES;
        ptx::mbar_wait(&empty[s], ((i / STAGES) & 1) ^ 1);
        ptx::mbar_arrive_expect_tx(&full[s], tx);
        uint8_t* sa = smem + s * stage_bytes;
        uint8_t* sb = sa + a_bytes;
        const int seg = seg0 + i % segs, img = img0 + i / segs;
        if (kConv3x3) {
          const int y0 = seg * p.HB, dy = tap / 3 - 1, dx = tap % 3 - 1;
          for (int j = 0; j < 2; ++j) ptx::tma_load_4d(sa + j * BOX_BYTES, &tmDz, &full[s], n0 + 64 * j, 0, y0, img);
          for (int j = 0; j < p.nboxes_b; ++j) ptx::tma_load_4d(sb + j * BOX_BYTES, &tmX, &full[s], k0 + 64 * j, dx, y0 + dy, img);
        } else {
          for (int j = 0; j < 2; ++j) ptx::tma_load_3d(sa + j * BOX_BYTES, &tmDz, &full[s], n0 + 64 * j, seg * SEG, img);
          for (int j = 0; j < p.nboxes_b; ++j) ptx::tma_load_3d(sb + j * BOX_BYTES, &tmX, &full[s], k0 + 64 * j, seg * SEG, img);
        }
      }
    }
  } else if (warp == 1) {
    // whole warp in the loop, one elected lane issues (see ptx::elect_one)
    // MN-major operands: LBO = distance between 64-channel boxes, SBO = distance between 8-pixel row groups
    const uint32_t idesc = ptx::umma_idesc_bf16(128, p.BNp) | (1u << 15) | (1u << 16);
    const uint32_t hi = static_cast<uint32_t>(1024 >> 4) | (1u << 14) | (2u << 29);
    const uint32_t smem_u = ptx::smem_u32(smem);
    for (int i = 0; i < nseg; ++i) {
      const int s = i % STAGES;
      ptx::mbar_wait(&full[s], (i / STAGES) & 1);
      ptx::tc_fence_after();
      const uint32_t sa = smem_u + s * stage_bytes, sb = sa + a_bytes;
      if (ptx::elect_one()) {
#pragma unroll
        for (int k = 0; k < SEG / 16; ++k) {
          const uint64_t adesc = ptx::umma_desc_mnmajor(sa + k * 2048, BOX_BYTES, hi);
          const uint64_t bdesc = ptx::umma_desc_mnmajor(sb + k * 2048, BOX_BYTES, hi);
          ptx::umma_bf16(tmem_base, adesc, bdesc, idesc, (i | k) != 0);
        }
        ptx::umma_commit(&empty[s]);
      }
      __syncwarp();
    }
    if (ptx::elect_one()) ptx::umma_commit(done);
    __syncwarp();
  } else {
    const int q = warp & 3;
    const int n = n0 + q * 32 + lane;
    ptx::mbar_wait(done, 0);
    ptx::tc_fence_after();
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    const float* sc = p.a_scale ? p.a_scale + static_cast<size_t>(img) * p.K : nullptr;
    for (int c0 = 0; c0 < p.BNp; c0 += 16) {
      uint32_t v[16];
      ptx::tmem_ld16(taddr + c0, v);
      ptx::tmem_ld_wait();
      if (n < p.N && nseg > 0) {
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          const int k = k0 + c0 + e;
          if (k < p.K) {
            float f = __uint_as_float(v[e]);
            if (sc) f *= __ldg(sc + k);
            atomicAdd(p.dw + (static_cast<size_t>(n) * p.K + k) * p.taps + tap, f);
          }
        }
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) { ptx::tc_fence_after(); ptx::tmem_dealloc(tmem_base, p.tmem_cols); }
}

}  // namespace

// returns MTG_ERR_UNSUPPORTED (without setting an error that matters) when the shape is better served by the CUDA-core kernel
int launch_wgrad_tc(const WgradArgs& a, int B, cudaStream_t st) {
  MTG_REQUIRE(a.dz && a.x && a.dw, MTG_ERR_ARG, "wgrad_tc: null pointer");
  MTG_REQUIRE(a.N % 8 == 0 && a.K % 8 == 0 && a.hw > 0 && B > 0 && static_cast<long long>(B) * a.hw == a.M, MTG_ERR_ARG,
              "wgrad_tc: bad geometry");
  const bool c3 = a.taps == 9;
  MTG_REQUIRE(!c3 || (a.W <= SEG && a.H * a.W == a.hw), MTG_ERR_UNSUPPORTED, "wgrad_tc: 3x3 needs W <= 64");
  WgTcP p{};
  p.N = a.N; p.K = a.K; p.taps = a.taps; p.dw = a.dw; p.a_scale = a.a_scale; p.H = a.H; p.W = a.W;
  p.nt_tiles = ceil_div(a.K, 256);
  p.BNp = static_cast<int>(align_up(ceil_div(a.K, p.nt_tiles), 16));
  p.nboxes_b = ceil_div(p.BNp, 64);
  p.mt_tiles = ceil_div(a.N, 128);
  uint32_t tm = 32;
  while (tm < static_cast<uint32_t>(p.BNp)) tm <<= 1;
  p.tmem_cols = tm;
  if (c3) {
    p.HB = SEG / a.W;
    if (p.HB > a.H) p.HB = a.H;
    p.rows_per_seg = p.HB * a.W;
    p.segs_per_img = ceil_div(a.H, p.HB);
  } else {
    p.rows_per_seg = SEG;
    p.segs_per_img = ceil_div(a.hw, SEG);
  }
  const int tiles = p.mt_tiles * p.nt_tiles * p.taps;
  // Every CTA adds its whole partial tile (128 x BNp fp32) to global memory with atomics, and shared memory allows one CTA
  // per SM: the pixel range is split for ONE wave of CTAs, not more (b14.expand at B=32: 512 CTAs / 10.5 M atomics before,
  // 128 CTAs / 2.6 M now).  Without a squeeze-excite gate a CTA may walk several images; with one the scale belongs to
  // the finished per-image accumulator, so those layers keep one image per CTA.
  int splits = 148 / tiles;
  if (splits < 1) splits = 1;
  int runs = 1;
  p.imgs_per_cta = 1;
  if (a.a_scale || splits >= B) {
    runs = splits / B;
    if (runs < 1) runs = 1;
    if (runs > p.segs_per_img) runs = p.segs_per_img;
  } else {
    p.imgs_per_cta = ceil_div(B, splits);
  }
  p.segs_per_cta = ceil_div(p.segs_per_img, runs);
  p.runs_per_img = ceil_div(p.segs_per_img, p.segs_per_cta);
  p.B = B;

  CUtensorMap tmDz, tmX;
  int rc;
  if (c3) {
    const unsigned long long dz_d[4] = {(unsigned long long)a.N, (unsigned long long)a.W, (unsigned long long)a.H, (unsigned long long)B};
    const unsigned long long dz_s[3] = {(unsigned long long)a.N * 2, (unsigned long long)a.W * a.N * 2, (unsigned long long)a.hw * a.N * 2};
    const unsigned box[4] = {64, (unsigned)a.W, (unsigned)p.HB, 1};
    rc = make_tma_map_bf16(&tmDz, a.dz, 4, dz_d, dz_s, box, 64);
    if (rc) return rc;
    const unsigned long long x_d[4] = {(unsigned long long)a.K, (unsigned long long)a.W, (unsigned long long)a.H, (unsigned long long)B};
    const unsigned long long x_s[3] = {(unsigned long long)a.K * 2, (unsigned long long)a.W * a.K * 2, (unsigned long long)a.hw * a.K * 2};
    rc = make_tma_map_bf16(&tmX, a.x, 4, x_d, x_s, box, 64);
    if (rc) return rc;
  } else {
    const unsigned long long dz_d[3] = {(unsigned long long)a.N, (unsigned long long)a.hw, (unsigned long long)B};
    const unsigned long long dz_s[2] = {(unsigned long long)a.N * 2, (unsigned long long)a.hw * a.N * 2};
    const unsigned box[3] = {64, SEG, 1};
    rc = make_tma_map_bf16(&tmDz, a.dz, 3, dz_d, dz_s, box, 64);
    if (rc) return rc;
    const unsigned long long x_d[3] = {(unsigned long long)a.K, (unsigned long long)a.hw, (unsigned long long)B};
    const unsigned long long x_s[2] = {(unsigned long long)a.K * 2, (unsigned long long)a.hw * a.K * 2};
    rc = make_tma_map_bf16(&tmX, a.x, 3, x_d, x_s, box, 64);
    if (rc) return rc;
  }
  const size_t smem = static_cast<size_t>(STAGES) * (2 + p.nboxes_b) * BOX_BYTES + 1024 + 128;
  dim3 grid(tiles, ceil_div(B, p.imgs_per_cta) * p.runs_per_img);
  if (c3) {
    static bool cfg = false;
    if (!cfg) { MTG_CUDA(cudaFuncSetAttribute(wgrad_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)); cfg = true; }
    wgrad_tc_kernel<true><<<grid, 192, smem, st>>>(tmDz, tmX, p);
  } else {
    static bool cfg = false;
    if (!cfg) { MTG_CUDA(cudaFuncSetAttribute(wgrad_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)); cfg = true; }
    wgrad_tc_kernel<false><<<grid, 192, smem, st>>>(tmDz, tmX, p);
  }
  MTG_LAUNCH_CHECK();
  return MTG_OK;
}

}  // namespace mtgseg
