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
// * warp roles: w0 = TMA producer, w1 = MMA issuer (+ TMEM owner), w2..9 = epilogue (4 TMEM lane quarters x
//   2 column halves: the epilogue is instruction-latency bound, so it gets two warps per scheduler),
//   w10..13 (only with kAScale) = squeeze-excite prologue that rescales the A tile in shared memory per
//   (image, channel) before the MMA reads it.
// * epilogue: TMEM -> registers -> folded BN / activation / residual -> bf16 -> 128B-swizzled staging slab
//   in shared memory -> ONE TMA store per 64-channel slab (double buffered); the residual tile arrives by
//   TMA as well, so the epilogue warps issue no per-thread global memory instruction at all and the
//   tensor map clips ragged rows / channels.
//
// Replaces, for this path, what the reference delegates to cuDNN/oneDNN: nn.Conv2d 1x1 in
// tv:models/mobilenetv3.py:71-80,101-105,179-187 and the head's 3x3 (train/model.py:110), with the
// eval-mode BatchNorm + activation (+ residual, tv:models/mobilenetv3.py:111-115) fused in the epilogue.
#include <cuda.h>
#include <stdio.h>
#include <stdlib.h>

#include <string.h>

#include <mutex>
#include <unordered_map>

#include "ops.h"
#include "ptx.cuh"

namespace mtgseg {

namespace {

constexpr int BM = 128;           // UMMA M
constexpr int MAX_STAGES = 8;
constexpr int BAR_BYTES = 384;               // mbarriers + TMEM slot
constexpr int MAX_RES = 8;                   // residual ring depth (mode 1)
constexpr int OUT_BUFS = 2;                  // double-buffered output staging
constexpr int SS_BYTES = 2 * 256 * 4;        // scale / shift of the current N tile (+ as much again for the BatchNorm statistics)

struct GemmKParams {
  int M, N, K;
  int BN, n_tiles, m_tiles;
  int num_kb, kb_per_tap, ksteps_last;
  int stages, tmem_cols, a_bytes;  // a_bytes: bytes one A TMA box delivers
  int kbox;                        // K elements per stage: 64 / 32 / 16 <-> 128B / 64B / 32B swizzled rows
  int b_res;                       // weights of this CTA's N tile stay resident in smem (few k-blocks): the ring carries A only
  uint32_t desc_hi;                // upper half of the smem matrix descriptors (SBO, version, swizzle mode)
  const float* scale;
  const float* shift;
  int act;
  const bf16* residual;
  bf16* out;
  const float* a_scale;
  int hw;
  int res_slabs;  // 0 or ceil(BN/obox)
  int res_bufs;   // 1: single buffer, reloaded after the epilogue has read it (long mainloops); 2: prefetched one tile ahead;
                  // 4..8 (mode 1, small tiles): a ring deep enough to cover the HBM latency of the residual stream as well
  int epi_mode;   // 1: per-warp epilogue (pointwise path): the two sets of 4 epilogue warps take alternate tiles, every warp stages and
                  //    TMA-stores its own 32-row sub-slab, no CTA-wide barrier; 0: all 8 warps share one slab at a time (3x3 / fallback)
  int wbufs;      // staging buffers per epilogue warp in mode 1 (1 or 2)
  int out_bytes;  // bytes of output staging in front of the residual buffers
  int obox;       // output / residual slab width in channels: 64 / 32 / 16 <-> 128B / 64B / 32B swizzled rows
  // multi-tap (3x3 / transposed-conv parity) geometry: A boxes {64 ch, W, HB rows, NB images} shifted by (dx, dy) per tap
  int B, H, W, HB, NB, h_tiles;
  signed char tap_dy[9], tap_dx[9];
  // training: per-channel sum / sum of squares of the (bf16-rounded) outputs, accumulated per CTA in shared memory and added to
  // stat[0..N) / stat[N..2N) with fp64 atomics when the CTA retires: the BatchNorm statistics come out of the conv epilogue
  double* stat;
  int ss_bytes;  // SS_BYTES, doubled when stat != nullptr
};

// acc[0..7] = x[0..7] * w[0..7] + acc[0..7] as four packed fp32x2 FMAs (FFMA2)
__device__ __forceinline__ void fma8_f2(float (&acc)[8], const float (&x)[8], const float (&w)[8]) {
#pragma unroll
  for (int j = 0; j < 8; j += 2) {
    uint64_t a, xv, wv;
    asm("mov.b64 %0, {%1,%2};" : "=l"(a) : "f"(acc[j]), "f"(acc[j + 1]));
    asm("mov.b64 %0, {%1,%2};" : "=l"(xv) : "f"(x[j]), "f"(x[j + 1]));
    asm("mov.b64 %0, {%1,%2};" : "=l"(wv) : "f"(w[j]), "f"(w[j + 1]));
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(a) : "l"(xv), "l"(wv));
    asm("mov.b64 {%0,%1}, %2;" : "=f"(acc[j]), "=f"(acc[j + 1]) : "l"(a));
  }
}

// z[0..7] *= t[0..7] as four packed fp32x2 multiplies (FMUL2)
__device__ __forceinline__ void mul8_f2(float (&z)[8], const float (&t)[8]) {
#pragma unroll
  for (int j = 0; j < 8; j += 2) {
    uint64_t a, b;
    asm("mov.b64 %0, {%1,%2};" : "=l"(a) : "f"(z[j]), "f"(z[j + 1]));
    asm("mov.b64 %0, {%1,%2};" : "=l"(b) : "f"(t[j]), "f"(t[j + 1]));
    asm("mul.rn.f32x2 %0, %0, %1;" : "+l"(a) : "l"(b));
    asm("mov.b64 {%0,%1}, %2;" : "=f"(z[j]), "=f"(z[j + 1]) : "l"(a));
  }
}

template <bool kConv3x3, bool kAScale>
__global__ void __launch_bounds__(kAScale ? 448 : 320, kAScale ? 1 : 2)
conv_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ CUtensorMap tmO, const __grid_constant__ CUtensorMap tmR, const GemmKParams p) {
  extern __shared__ uint8_t smem_raw[];
  pdl_trigger();  // the next kernel of the stream may be scheduled as soon as every CTA of this one is running
  // aligned by pointer arithmetic on the __shared__ base (not through uintptr_t): the compiler keeps the shared address space, so
  // the epilogue's scale / shift loads and staging stores are LDS / STS instead of generic LD.E / ST.E with 64-bit address math
  uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  const int S = p.stages;
  const int BK = p.kbox;
  const int A_STAGE_BYTES = BM * BK * 2;
  const int b_stage_bytes = p.BN * BK * 2;
  // B-stationary layout: [resident B: num_kb tiles][A ring] ; streaming layout: [A ring][B ring]
  const int b_res_bytes = p.b_res ? ((p.num_kb * b_stage_bytes + 1023) & ~1023) : 0;
  uint8_t* sBres = smem;
  uint8_t* sA = smem + b_res_bytes;
  uint8_t* sB = sA + S * A_STAGE_BYTES;
  uint8_t* sOut = smem + ((b_res_bytes + S * A_STAGE_BYTES + (p.b_res ? 0 : S * b_stage_bytes) + 1023) & ~1023);  // OUT_BUFS slabs (smem is 1024-aligned)
  const int SLAB_BYTES = BM * p.obox * 2;
  uint8_t* sRes = sOut + p.out_bytes;                  // res_bufs x res_slabs slabs
  float* sScale = reinterpret_cast<float*>(sRes + p.res_bufs * p.res_slabs * SLAB_BYTES);
  float* sShift = sScale + 256;
  // (only when p.stat) per-CTA BatchNorm partial sums: one slot per (row group, column), each owned by exactly one epilogue
  // thread, so they accumulate with plain adds in a fixed order: the statistics are run-to-run deterministic
  const int SG = 512 / p.obox;  // row groups: 128-row slab / (256 threads / column pairs) == 8 warps x row groups of a 32-row slab
  float* sSum = sShift + 256;   // [SG][BN]
  float* sSq = sSum + SG * p.BN;
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(sScale) + p.ss_bytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + MAX_STAGES;
  uint64_t* xform = bars + 2 * MAX_STAGES;
  uint64_t* tfull = bars + 3 * MAX_STAGES;
  uint64_t* tempty = tfull + 2;
  uint64_t* resbar = tempty + 2;  // [MAX_RES]: residual tile landed (indexed by ring slot; by epilogue set with a single buffer)
  uint64_t* bfull = resbar + MAX_RES;   // resident weights have landed
  uint64_t* res_free = bfull + 1; // [MAX_RES] (mode 1): the epilogue set has finished reading the residual tile in this slot
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(res_free + MAX_RES);

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
      ptx::mbar_init(&tempty[b], p.epi_mode ? 128 : 256);
    }
    for (int b = 0; b < MAX_RES; ++b) {
      ptx::mbar_init(&resbar[b], 1);
      ptx::mbar_init(&res_free[b], 128);
    }
    ptx::mbar_init(bfull, 1);
    ptx::fence_mbar_init();
  }
  if (warp == 0 && lane == 0) {
    ptx::tma_prefetch_desc(&tmA);
    ptx::tma_prefetch_desc(&tmB);
    ptx::tma_prefetch_desc(&tmO);
    if (p.res_slabs) ptx::tma_prefetch_desc(&tmR);
  }
  if (warp == 1) {
    ptx::tmem_alloc(tmem_slot, p.tmem_cols);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  pdl_wait();  // barrier init, TMEM allocation and descriptor prefetch above overlapped the predecessor's tail; its output is read from here on
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t buf_stride = p.tmem_cols >> 1;
  const int total_tiles = p.m_tiles * p.n_tiles;

  if (warp == 0) {
    // =============================== TMA producer ===============================
    if (lane == 0) {
      uint32_t it = 0;
      const uint32_t stage_bytes = p.a_bytes + (p.b_res ? 0 : b_stage_bytes);
      if (p.b_res && static_cast<int>(blockIdx.x) < total_tiles) {
        // gridDim.x is a multiple of n_tiles, so this CTA's N tile never changes: fetch its weights once
        const int n_tile = blockIdx.x % p.n_tiles;
        ptx::mbar_arrive_expect_tx(bfull, p.num_kb * b_stage_bytes);
        for (int kb = 0; kb < p.num_kb; ++kb)
          ptx::tma_load_2d(sBres + kb * b_stage_bytes, &tmB, bfull, kb * BK, n_tile * p.BN);
      }
      // mode 1: the producer also fetches the residual tiles (the epilogue sets run decoupled, there is no leader thread).
      // Residual loads never block the A/B ring: they are issued from `pump` whenever their buffer has been released.
      const int r_slabs = p.res_slabs, r_slab_bytes = BM * p.obox * 2;
      uint32_t tcr = 0;        // residual tiles issued so far (per-CTA tile counter)
      int rtile = blockIdx.x;  // tile index of the next residual to issue
      // Barriers are indexed by the epilogue SET (= tile-counter parity), not by the buffer: each barrier then has exactly
      // one sequential consumer, so the usual (use count & 1) parity is unambiguous even when one buffer serves both sets.
      auto pump = [&](uint32_t upto) {  // issue pending residual loads of tiles with counter <= upto whose buffer is free
        while (rtile < total_tiles && tcr <= upto) {
          // ring slot and barrier of residual tile t: slot t % res_bufs (an even ring: slot parity == epilogue set, so every
          // barrier has one sequential consumer); with a single buffer the barriers are indexed by the set instead
          const uint32_t rb = p.res_bufs >= 2 ? tcr % p.res_bufs : 0;
          const uint32_t bi = p.res_bufs >= 2 ? rb : (tcr & 1);
          // the slot was last read for tile tcr - res_bufs (use number (tcr - res_bufs) / res_bufs of the slot; single buffer:
          // by set (tcr - 1) & 1 as that set's ((tcr - 1) >> 1)-th tile)
          if (tcr >= static_cast<uint32_t>(p.res_bufs)) {
            const uint32_t prev = tcr - p.res_bufs;
            const bool ok = p.res_bufs >= 2 ? ptx::mbar_try_wait(&res_free[rb], (prev / p.res_bufs) & 1)
                                            : ptx::mbar_try_wait(&res_free[prev & 1], (prev >> 1) & 1);
            if (!ok) return false;
          }
          const int mt = rtile / p.n_tiles, nt = rtile - mt * p.n_tiles;
          ptx::mbar_arrive_expect_tx(&resbar[bi], r_slabs * r_slab_bytes);
          for (int sl = 0; sl < r_slabs; ++sl)
            ptx::tma_load_2d(sRes + (rb * r_slabs + sl) * r_slab_bytes, &tmR, &resbar[bi], nt * p.BN + sl * p.obox, mt * BM);
          ++tcr;
          rtile += gridDim.x;
        }
        return true;
      };
      const bool pump_res = p.epi_mode == 1 && p.res_slabs > 0;
      uint32_t tcp = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++tcp) {
        const int m_tile = tile / p.n_tiles, n_tile = tile - m_tile * p.n_tiles;
        int n0 = 0, y0 = 0;
        if (kConv3x3) {
          n0 = (m_tile / p.h_tiles) * p.NB;
          y0 = (m_tile % p.h_tiles) * p.HB;
        }
        for (int kb = 0; kb < p.num_kb; ++kb, ++it) {
          const int s = it % S;
          const uint32_t ph = (it / S) & 1;
          if (pump_res) {
            pump(tcp);
            while (!ptx::mbar_try_wait(&empty[s], ph ^ 1)) pump(tcp);
          } else {
            ptx::mbar_wait(&empty[s], ph ^ 1);
          }
          ptx::mbar_arrive_expect_tx(&full[s], stage_bytes);
          if (kConv3x3) {
            const int tap = kb / p.kb_per_tap, kc = kb - tap * p.kb_per_tap;
            ptx::tma_load_4d(sA + s * A_STAGE_BYTES, &tmA, &full[s], kc * BK, p.tap_dx[tap], y0 + p.tap_dy[tap], n0);
            ptx::tma_load_2d(sB + s * b_stage_bytes, &tmB, &full[s], tap * p.K + kc * BK, n_tile * p.BN);
          } else {
            ptx::tma_load_2d(sA + s * A_STAGE_BYTES, &tmA, &full[s], kb * BK, m_tile * BM);
            if (!p.b_res) ptx::tma_load_2d(sB + s * b_stage_bytes, &tmB, &full[s], kb * BK, n_tile * p.BN);
          }
        }
      }
      if (pump_res) {  // drain: the remaining residual tiles, now with blocking waits (bounded like every mbarrier wait here)
        const long long t0 = clock64();
        while (!pump(0xffffffffu)) {
          if (clock64() - t0 > 8000000000LL) __trap();
        }
      }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer ===============================
    // The whole warp walks the loop in warp-uniform control flow and one elected lane issues: stage indices and descriptors
    // then live in uniform registers.  (Inside an `if (lane == 0)` region the compiler wraps every UTCHMMA in an ELECT / R2UR
    // loop: measured ~130 cycles of issue per MMA, which bounded every tensor-heavy layer at ~35 % of the tensor pipe.)
    {
      const uint32_t idesc = ptx::umma_idesc_bf16(BM, p.BN);
      const uint32_t sA_u = ptx::smem_u32(sA), sB_u = ptx::smem_u32(sB), sBres_u = ptx::smem_u32(sBres);
      uint32_t it = 0, tc = 0;
      if (p.b_res && static_cast<int>(blockIdx.x) < total_tiles) ptx::mbar_wait(bfull, 0);
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
          const uint64_t adesc = ptx::umma_desc_kmajor(sA_u + s * A_STAGE_BYTES, p.desc_hi);
          const uint64_t bdesc = ptx::umma_desc_kmajor(p.b_res ? sBres_u + kb * b_stage_bytes : sB_u + s * b_stage_bytes, p.desc_hi);
          if (ptx::elect_one()) {
#pragma unroll
            for (int k = 0; k < 4; ++k)  // +32 B along K inside the swizzle row == +2 in the address field
              if (k < ksteps) ptx::umma_bf16(d, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
            ptx::umma_commit(&empty[s]);  // frees the smem stage once these MMAs have read it
          }
          __syncwarp();
        }
        if (ptx::elect_one()) ptx::umma_commit(&tfull[buf]);  // accumulator complete -> epilogue
        __syncwarp();
      }
    }
  } else if (warp < 10) {
    // =============================== epilogue (8 warps) ===============================
    const int q = warp & 3;                  // TMEM lane quarter this warp may access
    const int wset = (warp - 2) >> 2;        // mode 0: which 32 of the 64 slab columns; mode 1: which tiles (alternate)
    const int r = q * 32 + lane;             // accumulator row of this thread
    const int et = threadIdx.x - 64;         // 0..255 among the epilogue threads
    const bool leader = et == 0;
    const int OB = p.obox, opitch = OB * 2;  // slab columns, slab row pitch in bytes
    const int slabs = (p.BN + OB - 1) / OB;
    // swizzle XOR term of this row: 128B rows: row % 8 ; 64B rows: (row / 2) % 4 ; 32B rows: (row / 4) % 2
    const int rx = OB == 64 ? (r & 7) : (OB == 32 ? ((r >> 1) & 3) : ((r >> 2) & 1));
    // 32 accumulator columns [sl * OB + half * 32, +32) of this thread's row: one TMEM round trip, folded BN, activation,
    // residual, bf16, into the swizzled staging row `srow` (rrow: the same row of the residual slab)
    auto convert_half = [&](uint32_t taddr, int sl, int half, int cols, uint8_t* srow, const uint8_t* rrow) {
      uint32_t v[2][16];
      ptx::tmem_ld16(taddr + sl * OB + half * 32, v[0]);
      if (half * 32 + 16 < cols) ptx::tmem_ld16(taddr + sl * OB + half * 32 + 16, v[1]);
      ptx::tmem_ld_wait();
      // the activation is uniform over the launch: dispatch once per call, not once per element
      auto convert = [&](auto actf) {
#pragma unroll
        for (int cc = 0; cc < 2; ++cc) {
          if (half * 32 + cc * 16 < cols) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const int chunk = half * 4 + cc * 2 + h;       // 16-byte chunk inside the slab row
              const int c = sl * OB + chunk * 8;             // column inside the N tile
              const int phys = (chunk ^ rx) * 16;            // TMA swizzle of the staging slab
              const float4 s0 = *reinterpret_cast<const float4*>(sScale + c), s1 = *reinterpret_cast<const float4*>(sScale + c + 4);
              const float4 h0 = *reinterpret_cast<const float4*>(sShift + c), h1 = *reinterpret_cast<const float4*>(sShift + c + 4);
              const float sc[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
              float f[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
              float a[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) a[e] = __uint_as_float(v[cc][h * 8 + e]);
              fma8_f2(f, a, sc);  // f = acc * scale + shift, four packed fp32x2 FMAs
              actf(f);
              if (p.res_slabs) {
                float rf[8];
                unpack8(*reinterpret_cast<const uint4*>(rrow + phys), rf);
#pragma unroll
                for (int e = 0; e < 8; ++e) f[e] += rf[e];
              }
              *reinterpret_cast<uint4*>(srow + phys) = pack8(f);
            }
          }
        }
      };
      // hardswish as z * sat(z/6 + 1/2): one FFMA.SAT per element + one packed FMUL2 per pair instead of add / max / min / mul / mul
      if (p.act == ACT_HSWISH) convert([](float (&z)[8]) {
        float t[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) t[e] = __saturatef(fmaf(z[e], 1.f / 6.f, 0.5f));
        mul8_f2(z, t);
      });
      else if (p.act == ACT_RELU) convert([](float (&z)[8]) {
#pragma unroll
        for (int e = 0; e < 8; ++e) z[e] = fmaxf(z[e], 0.f);
      });
      else if (p.act == ACT_NONE) convert([](float (&)[8]) {});
      else convert([&](float (&z)[8]) {
#pragma unroll
        for (int e = 0; e < 8; ++e) z[e] = apply_act(z[e], p.act);
      });
    };
    // ---- BatchNorm statistics of a staged slab (training): column sums of the bf16 values the TMA store writes --------------
    if (p.stat)
      for (int i = et; i < 2 * SG * p.BN; i += 256) sSum[i] = 0.f;  // visible after the first named barrier below (both modes have one)
    // `nthreads` consecutive threads (index ti) share a slab of `nrows` rows (box-local row index = swizzle row): a thread owns
    // one bf16 column pair and a group of rows; lanes of a warp read one 128-byte row segment -> conflict free
    auto slab_stats = [&](const uint8_t* sbuf, int sl, int cols, int ti, int nthreads, int nrows, int row_limit, int group_base) {
      const int pairs = OB >> 1, groups = nthreads / pairs, rpg = nrows / groups;
      const int cp = ti % pairs, grp = ti / pairs;
      if (cp * 2 >= cols) return;
      float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
      const int chunk = cp >> 2, within = (cp & 3) * 4;
      for (int i = 0; i < rpg; ++i) {
        const int row = grp * rpg + i;
        if (row >= row_limit) break;
        const int rxx = OB == 64 ? (row & 7) : (OB == 32 ? ((row >> 1) & 3) : ((row >> 2) & 1));
        const uint32_t w = *reinterpret_cast<const uint32_t*>(sbuf + row * opitch + ((chunk ^ rxx) << 4) + within);
        const float a = __uint_as_float(w << 16), b = __uint_as_float(w & 0xFFFF0000u);
        s0 += a; s1 += b;
        q0 = fmaf(a, a, q0); q1 = fmaf(b, b, q1);
      }
      const int slot = (group_base + grp) * p.BN + sl * OB + cp * 2;  // this thread's own slot: no atomics
      sSum[slot] += s0; sSum[slot + 1] += s1;
      sSq[slot] += q0; sSq[slot + 1] += q1;
    };
    // all 256 epilogue threads: add this CTA's sums of N tile `n_tile` to the global fp64 accumulators, clear the local ones
    auto flush_stats = [&](int n_tile) {
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (et < p.BN) {
        const int n = n_tile * p.BN + et;
        float a = 0.f, b = 0.f;
        for (int g = 0; g < SG; ++g) {
          a += sSum[g * p.BN + et]; b += sSq[g * p.BN + et];
          sSum[g * p.BN + et] = 0.f; sSq[g * p.BN + et] = 0.f;
        }
        if (n < p.N) {
          atomicAdd(p.stat + n, static_cast<double>(a));
          atomicAdd(p.stat + p.N + n, static_cast<double>(b));
        }
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
    };
    if (!kConv3x3 && p.epi_mode == 1) {
      // ---- mode 1: decoupled warps.  Set `wset` owns the tiles with (tile counter & 1) == wset and the TMEM buffer of the
      // same index, so two tiles are in their epilogue at once; each warp converts its 32 rows of every slab into its own
      // staging buffer and stores it with its own TMA instruction ({OB channels, 32 rows} box): only __syncwarp inside.
      const int WSLAB = 32 * opitch;
      uint8_t* wbuf = sOut + (warp - 2) * p.wbufs * WSLAB;
      {  // folded-BN constants of this CTA's N tile (fixed for its whole life: one N tile, or resident weights)
        const int n_tile = blockIdx.x % p.n_tiles;
        for (int c = et; c < p.BN; c += 256) {
          const int n = n_tile * p.BN + c;
          sScale[c] = (p.scale && n < p.N) ? __ldg(p.scale + n) : 1.f;
          sShift[c] = (p.shift && n < p.N) ? __ldg(p.shift + n) : 0.f;
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
      }
      uint32_t tc = 0, wstore = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++tc) {
        if (static_cast<int>(tc & 1) != wset) continue;
        const int m_tile = tile / p.n_tiles, n_tile = tile - m_tile * p.n_tiles;
        const uint32_t buf = tc & 1, aph = (tc >> 1) & 1;
        const uint32_t rb = p.res_bufs >= 2 ? tc % p.res_bufs : 0;  // residual ring slot (see the producer)
        const uint32_t rbi = p.res_bufs >= 2 ? rb : static_cast<uint32_t>(wset);
        ptx::mbar_wait(&tfull[buf], aph);
        ptx::tc_fence_after();
        if (p.res_slabs) ptx::mbar_wait(&resbar[rbi], p.res_bufs >= 2 ? ((tc / p.res_bufs) & 1) : ((tc >> 1) & 1));
        const uint32_t taddr = tmem_base + buf * buf_stride + (static_cast<uint32_t>(q * 32) << 16);
        for (int sl = 0; sl < slabs; ++sl, ++wstore) {
          uint8_t* sbuf = wbuf + (p.wbufs == 2 ? (wstore & 1) : 0) * WSLAB;
          if (lane == 0) {  // the store that used this buffer before must have finished reading it
            if (p.wbufs == 2) ptx::bulk_wait_read<1>();
            else ptx::bulk_wait_read<0>();
          }
          __syncwarp();
          const int cols = min(OB, p.BN - sl * OB);  // multiple of 16
          uint8_t* srow = sbuf + lane * opitch;
          const uint8_t* rrow = sRes + (rb * slabs + sl) * SLAB_BYTES + r * opitch;
          convert_half(taddr, sl, 0, cols, srow, rrow);
          if (32 < cols) convert_half(taddr, sl, 1, cols, srow, rrow);
          ptx::fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            ptx::tma_store_2d(&tmO, sbuf, n_tile * p.BN + sl * OB, m_tile * BM + q * 32);
            ptx::bulk_commit();
          }
          if (p.stat) slab_stats(sbuf, sl, cols, lane, 32, 32, 32, (warp - 2) * (64 / OB));  // (the next use of sbuf starts with a __syncwarp)
        }
        ptx::tc_fence_before();
        ptx::mbar_arrive(&tempty[buf]);
        if (p.res_slabs) ptx::mbar_arrive(&res_free[rbi]);
      }
      if (p.stat) flush_stats(blockIdx.x % p.n_tiles);
      if (lane == 0) ptx::bulk_wait_all();  // smem must stay valid until the last store has read it
    } else {
      // ---- mode 0: all 8 warps work on one slab at a time (4 lane quarters x 2 column halves), one leader thread stores
      const int half = wset;
      uint32_t tc = 0, store_no = 0;
      int cur_ntile = -1;
      auto load_residual = [&](int t, uint32_t rb) {  // leader only: residual tile of `t` -> buffer rb by TMA
        const int mt = t / p.n_tiles, nt = t - mt * p.n_tiles;
        ptx::mbar_arrive_expect_tx(&resbar[rb], slabs * SLAB_BYTES);
        for (int sl = 0; sl < slabs; ++sl)
          ptx::tma_load_2d(sRes + (rb * slabs + sl) * SLAB_BYTES, &tmR, &resbar[rb], nt * p.BN + sl * OB, mt * BM);
      };
      if (p.res_slabs && leader && static_cast<int>(blockIdx.x) < total_tiles) load_residual(blockIdx.x, 0);
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++tc) {
        const int m_tile = tile / p.n_tiles, n_tile = tile - m_tile * p.n_tiles;
        const uint32_t buf = tc & 1, aph = (tc >> 1) & 1;
        int oc1 = m_tile * BM, oc2 = 0, oc3 = 0;  // output coordinates after the channel coordinate
        if (kConv3x3) {
          oc1 = 0;
          oc2 = (m_tile % p.h_tiles) * p.HB;
          oc3 = (m_tile / p.h_tiles) * p.NB;
        }
        const uint32_t rb = p.res_bufs == 2 ? (tc & 1) : 0;
        // prefetch the NEXT tile's residual into the other buffer (its last readers finished before the final
        // bar.sync of the previous iteration)
        if (p.res_slabs && p.res_bufs == 2 && leader && tile + static_cast<int>(gridDim.x) < total_tiles) load_residual(tile + gridDim.x, rb ^ 1);
        if (n_tile != cur_ntile) {  // folded-BN constants of this N tile -> smem (visible after the first bar.sync below)
          if (p.stat && cur_ntile >= 0) flush_stats(cur_ntile);
          cur_ntile = n_tile;
          for (int c = et; c < p.BN; c += 256) {
            const int n = n_tile * p.BN + c;
            sScale[c] = (p.scale && n < p.N) ? __ldg(p.scale + n) : 1.f;
            sShift[c] = (p.shift && n < p.N) ? __ldg(p.shift + n) : 0.f;
          }
        }
        ptx::mbar_wait(&tfull[buf], aph);
        ptx::tc_fence_after();
        if (p.res_slabs) ptx::mbar_wait(&resbar[rb], p.res_bufs == 2 ? ((tc >> 1) & 1) : (tc & 1));
        const uint32_t taddr = tmem_base + buf * buf_stride + (static_cast<uint32_t>(q * 32) << 16);
        for (int sl = 0; sl < slabs; ++sl, ++store_no) {
          uint8_t* sbuf = sOut + (store_no & 1) * SLAB_BYTES;
          // the TMA store that used this buffer two slabs ago must have finished reading it
          if (leader) ptx::bulk_wait_read<OUT_BUFS - 1>();
          asm volatile("bar.sync 1, 256;" ::: "memory");
          const int cols = min(OB, p.BN - sl * OB);  // multiple of 16
          if (half * 32 < cols)
            convert_half(taddr, sl, half, cols, sbuf + r * opitch, sRes + (rb * slabs + sl) * SLAB_BYTES + r * opitch);
          ptx::fence_proxy_async_smem();
          asm volatile("bar.sync 1, 256;" ::: "memory");
          if (leader) {
            const int c0 = n_tile * p.BN + sl * OB;
            if (kConv3x3) ptx::tma_store_4d(&tmO, sbuf, c0, oc1, oc2, oc3);
            else ptx::tma_store_2d(&tmO, sbuf, c0, oc1);
            ptx::bulk_commit();
          }
          // (the slab is reused two slabs later, behind the next slab's first barrier.)  3x3: accumulator rows beyond the TMA box
          // were computed from stale shared memory and are never stored: they must not be counted either
          if (p.stat) slab_stats(sbuf, sl, cols, et, 256, BM, kConv3x3 ? p.NB * p.HB * p.W : BM, 0);
        }
        // single residual buffer: every epilogue thread passed the last bar.sync after its final read of the tile (and
        // executed fence.proxy.async before it), so the buffer may be refilled now; the next mainloop hides the load
        if (p.res_slabs && p.res_bufs == 1 && leader && tile + static_cast<int>(gridDim.x) < total_tiles) load_residual(tile + gridDim.x, 0);
        ptx::tc_fence_before();
        ptx::mbar_arrive(&tempty[buf]);
      }
      if (p.stat && cur_ntile >= 0) flush_stats(cur_ntile);
      if (leader) ptx::bulk_wait_all();  // smem must stay valid until the last store has read it
    }
  } else if (kAScale) {
    // =============================== squeeze-excite prologue on the A tile ===============================
    const int t = threadIdx.x - 320;
    const int chunk = t & 7;  // physical 16-byte chunk inside the 128-byte row
    // this thread's 16-byte chunk holds the same 8 logical channels in every row it touches (rows differ by 16, the
    // swizzle only looks at row % 8), and a 128-row tile spans at most two images when hw >= 128: the (at most two)
    // gate vectors of a stage are fetched ONE STAGE AHEAD into registers, so their global-load latency is hidden behind
    // the previous stage instead of sitting between "A tile landed" and "MMA may start"
    const int kofs = (chunk ^ ((t >> 3) & 7)) << 3;
    const bool fast = p.hw >= BM;
    float na[8], nb[8];
    int nsplit = 0;  // rows are indexed in 32 bits here: M is an int and a tile starts below M
    auto fetch = [&](int tile, int kb) {
      const int k = kb * BK + kofs;
      const int m0 = (tile / p.n_tiles) * BM;
      const int img0 = m0 / p.hw;
      nsplit = (img0 + 1) * p.hw;  // first row index of the next image
      if (fast && k < p.K) {
        const float* sp = p.a_scale + static_cast<size_t>(img0) * p.K + k;
        const float4 a0 = __ldg(reinterpret_cast<const float4*>(sp)), a1 = __ldg(reinterpret_cast<const float4*>(sp + 4));
        na[0] = a0.x; na[1] = a0.y; na[2] = a0.z; na[3] = a0.w; na[4] = a1.x; na[5] = a1.y; na[6] = a1.z; na[7] = a1.w;
        const bool two = nsplit < m0 + BM && nsplit < p.M;
        const float* sq = two ? sp + p.K : sp;
        const float4 b0 = __ldg(reinterpret_cast<const float4*>(sq)), b1 = __ldg(reinterpret_cast<const float4*>(sq + 4));
        nb[0] = b0.x; nb[1] = b0.y; nb[2] = b0.z; nb[3] = b0.w; nb[4] = b1.x; nb[5] = b1.y; nb[6] = b1.z; nb[7] = b1.w;
      }
    };
    if (static_cast<int>(blockIdx.x) < total_tiles) fetch(blockIdx.x, 0);
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int m_tile = tile / p.n_tiles;
      const int m0 = m_tile * BM;
      for (int kb = 0; kb < p.num_kb; ++kb, ++it) {
        const int s = it % S;
        const uint32_t ph = (it / S) & 1;
        float sa[8], sb[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) { sa[e] = na[e]; sb[e] = nb[e]; }
        const int split = nsplit;
        {
          int ntile = tile, nkb = kb + 1;
          if (nkb == p.num_kb) { nkb = 0; ntile = tile + gridDim.x; }
          if (ntile < total_tiles) fetch(ntile, nkb);
        }
        ptx::mbar_wait(&full[s], ph);
        uint8_t* base = sA + s * A_STAGE_BYTES;
        const int k = kb * BK + kofs;
        if (k < p.K && !fast) {  // tiny feature maps (tests): a tile may span many images, fetch the gate per row
          for (int j = 0; j < 8; ++j) {
            const int row = (t >> 3) + 16 * j;
            const int m = m0 + row;
            if (m < p.M) {
              const float* sp = p.a_scale + static_cast<size_t>(m / p.hw) * p.K + k;
              uint4* ptr = reinterpret_cast<uint4*>(base + row * 128 + chunk * 16);
              float f[8];
              unpack8(*ptr, f);
#pragma unroll
              for (int e = 0; e < 8; ++e) f[e] *= __ldg(sp + e);
              *ptr = pack8(f);
            }
          }
        } else if (k < p.K) {
          // all eight 16-byte chunks of this thread are loaded BEFORE any is stored back: written as load / scale / store
          // per row the compiler must keep every shared-memory load behind the previous row's store (same buffer), and
          // the stage became eight dependent LDS -> ALU -> STS chains (measured: +35 us on a 52 us layer)
          uint4 qv[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int row = (t >> 3) + 16 * j;
            qv[j] = *reinterpret_cast<const uint4*>(base + row * 128 + chunk * 16);
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int row = (t >> 3) + 16 * j;
            const bool second = m0 + row >= split;
            float f[8];
            unpack8(qv[j], f);
#pragma unroll
            for (int e = 0; e < 8; ++e) f[e] *= second ? sb[e] : sa[e];
            qv[j] = pack8(f);
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int row = (t >> 3) + 16 * j;
            if (m0 + row < p.M) *reinterpret_cast<uint4*>(base + row * 128 + chunk * 16) = qv[j];
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

// bf16 tensor map, zero OOB fill; swizzle by kbox: 64 -> 128 B, 32 -> 64 B, 16 -> 32 B, 0 -> none (dense box, used by the
// depthwise tiles).  dims/strides innermost first; strides in bytes for dims 1..
// Encoded maps are cached per thread, keyed on everything that goes into them (base pointer, geometry, box, swizzle): a training
// step needs ~450 maps over ~150 distinct (tensor, tiling) pairs that repeat every step (the workspace layout is fixed), and
// cuTensorMapEncodeTiled was a visible share of the 4 ms the host needs to enqueue an eager step.
struct MapKey {
  const void* base; int rank, kbox; uint64_t dims[5]; uint64_t strides[4]; uint32_t box[5];
  bool operator==(const MapKey& o) const { return memcmp(this, &o, sizeof(MapKey)) == 0; }
};
struct MapKeyHash {
  size_t operator()(const MapKey& k) const {
    const uint64_t* w = reinterpret_cast<const uint64_t*>(&k);
    uint64_t h = 1469598103934665603ull;
    for (size_t i = 0; i < sizeof(MapKey) / 8; ++i) { h ^= w[i]; h *= 1099511628211ull; }
    return static_cast<size_t>(h);
  }
};

int make_map(CUtensorMap* map, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
             const uint32_t* box, int kbox = 64) {
  static thread_local std::unordered_map<MapKey, CUtensorMap, MapKeyHash> cache;
  MapKey key;
  memset(&key, 0, sizeof(key));
  key.base = base; key.rank = rank; key.kbox = kbox;
  for (int i = 0; i < rank; ++i) {
    key.dims[i] = dims[i];
    key.box[i] = box[i];
    if (i > 0) key.strides[i - 1] = strides_bytes[i - 1];
  }
  auto hit = cache.find(key);
  if (hit != cache.end()) {
    *map = hit->second;
    return MTG_OK;
  }
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
                  CU_TENSOR_MAP_INTERLEAVE_NONE,
                  kbox == 64 ? CU_TENSOR_MAP_SWIZZLE_128B
                             : (kbox == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : (kbox == 16 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_NONE)),
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  MTG_REQUIRE(r == CUDA_SUCCESS, MTG_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d (rank %d, dim0 %llu)",
              static_cast<int>(r), rank, static_cast<unsigned long long>(dims[0]));
  if (cache.size() >= 4096) cache.clear();  // bounded: varying batch shapes / reallocated workspaces must not grow it forever
  cache.emplace(key, *map);
  return MTG_OK;
}

}  // namespace

// shared with wgrad_tc.cu and dwconv.cu
int make_tma_map_bf16(CUtensorMap* map, const void* base, int rank, const unsigned long long* dims,
                      const unsigned long long* strides_bytes, const unsigned* box, int kbox) {
  uint64_t d[5], s[4];
  uint32_t b[5];
  for (int i = 0; i < rank; ++i) { d[i] = dims[i]; b[i] = box[i]; if (i > 0) s[i - 1] = strides_bytes[i - 1]; }
  return make_map(map, base, rank, d, s, b, kbox);
}

namespace {

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
int launch_variant(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmO, const CUtensorMap& tmR,
                   const GemmKParams& kp, size_t need, cudaStream_t st) {
  static bool configured = false;  // per template instantiation
  if (!configured) {
    MTG_CUDA(cudaFuncSetAttribute(conv_gemm_kernel<C3, AS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    configured = true;
  }
  const int threads = AS ? 448 : 320;
  // co-resident CTAs per SM from the kernel's own resource use: registers (allocated in units of 8 per thread), shared
  // memory (+1 KB reserved per CTA), threads, and TMEM columns (all allocations must fit the 512 columns of an SM).
  // cudaOccupancyMaxActiveBlocksPerMultiprocessor is NOT used: on this driver it answers 1 for this kernel for every
  // block size / shared-memory request (measured on a B200, even 256 threads and 0 bytes), while two CTAs do run
  // concurrently (29-31 % warps active under ncu with a 296-CTA grid).
  static int regs_per_thread = 0;  // per template instantiation
  if (regs_per_thread == 0) {
    cudaFuncAttributes fa{};
    MTG_CUDA(cudaFuncGetAttributes(&fa, conv_gemm_kernel<C3, AS>));
    regs_per_thread = fa.numRegs > 0 ? (fa.numRegs + 7) / 8 * 8 : 128;
  }
  int per_sm = 65536 / (regs_per_thread * threads);
  const int by_smem = static_cast<int>((227 * 1024) / (need + 1024));
  if (per_sm > by_smem) per_sm = by_smem;
  if (per_sm > 2048 / threads) per_sm = 2048 / threads;
  const int occ_raw = per_sm;
  if (per_sm > 512 / kp.tmem_cols) per_sm = 512 / kp.tmem_cols;
  if (per_sm > 4) per_sm = 4;
  if (per_sm < 1) per_sm = 1;
  // pad the request so that never more than per_sm CTAs share an SM (a persistent CTA that had to wait for TMEM
  // or for a slot would serialise behind a whole tile list)
  size_t smem = need;
  const size_t floor_for_cap = (227 * 1024) / (per_sm + 1) + 1;
  if (smem < floor_for_cap) smem = floor_for_cap;
  const long long total_tiles = static_cast<long long>(kp.m_tiles) * kp.n_tiles;
  int grid = num_sms() * per_sm;
  if (grid > total_tiles) grid = static_cast<int>(total_tiles);
  if (kp.b_res || kp.stat) {  // resident weights / per-CTA statistics: a CTA keeps one N tile for its whole life -> grid is a multiple of n_tiles
    grid = grid / kp.n_tiles * kp.n_tiles;
    if (grid < kp.n_tiles) grid = kp.n_tiles;
  }
  static const bool debug = getenv("MTGSEG_GEMM_DEBUG") != nullptr;
  if (debug) {
    cudaFuncAttributes fa{};
    cudaFuncGetAttributes(&fa, conv_gemm_kernel<C3, AS>);
    fprintf(stderr, "[conv_gemm<%d,%d>] occ_raw=%d regs=%d static_smem=%zu local=%zu max_threads=%d max_dyn=%d\n", int(C3), int(AS), occ_raw,
            fa.numRegs, fa.sharedSizeBytes, fa.localSizeBytes, fa.maxThreadsPerBlock, fa.maxDynamicSharedSizeBytes);
  }
  if (debug)
    fprintf(stderr, "[conv_gemm<%d,%d>] M=%d N=%d K=%d BN=%d kbox=%d obox=%d num_kb=%d stages=%d b_res=%d res_slabs=%d res_bufs=%d epi=%d wbufs=%d tmem=%d need=%zu smem=%zu per_sm=%d grid=%d tiles=%lld\n",
            int(C3), int(AS), kp.M, kp.N, kp.K, kp.BN, kp.kbox, kp.obox, kp.num_kb, kp.stages, kp.b_res, kp.res_slabs, kp.res_bufs, kp.epi_mode, kp.wbufs, kp.tmem_cols, need, smem,
            per_sm, grid, total_tiles);
  MTG_CUDA(launch_pdl(conv_gemm_kernel<C3, AS>, dim3(grid), dim3(threads), smem, st, tmA, tmB, tmO, tmR, kp));
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
  if (g.conv3x3) {  // plain 3x3 on small feature maps: the haloed-tile kernel (conv3_tc.cu)
    const int rc = launch_conv3x3_halo(g, st);
    if (rc <= 0) return rc;
  }

  GemmKParams kp{};
  kp.M = g.M; kp.N = g.N; kp.K = g.K;
  // (capping the N tile at 128 columns for K <= 256 so that two CTAs share an SM, 2 x 128 TMEM columns each, was measured and
  // rejected: b7.expand 52 -> 44 us but b14-16.expand 60 -> 73 us, the A tile is then fetched eight times)
  if (g.N <= 256) {
    kp.n_tiles = 1;
    kp.BN = static_cast<int>(align_up(g.N, 16));
  } else {
    // several N tiles: BN must be a multiple of the 64-channel store slab (a slab may never spill into the next
    // N tile); take the BN with the fewest padded columns, the largest on ties (fewest re-reads of A)
    int best = 256, best_cols = 1 << 30;
    for (int bn = 256; bn >= 64; bn -= 64) {
      const int cols = ceil_div(g.N, bn) * bn;
      if (cols < best_cols) { best_cols = cols; best = bn; }
    }
    kp.BN = best;
    kp.n_tiles = ceil_div(g.N, best);
  }
  // narrow layers (K <= 32) use narrower swizzled rows so that a stage is 6-12 KB instead of 24 KB: the ring can then
  // be deep enough to cover HBM latency (these layers move 4 KB of A per tile)
  const int BK = (!g.conv3x3 && !g.a_scale && g.K <= 16) ? 16 : ((!g.conv3x3 && !g.a_scale && g.K <= 32) ? 32 : 64);
  const int A_STAGE_BYTES = BM * BK * 2;
  kp.kbox = BK;
  {
    const uint32_t sbo = static_cast<uint32_t>(8 * BK * 2) >> 4;             // 8 rows of the swizzle atom
    const uint32_t layout = BK == 64 ? 2u : (BK == 32 ? 4u : 6u);            // SWIZZLE_128B / 64B / 32B
    kp.desc_hi = sbo | (1u << 14) | (layout << 29);                          // bits 32-45 SBO, 46-48 version 1, 61-63 layout
  }
  kp.kb_per_tap = ceil_div(g.K, BK);
  const int ntaps = g.conv3x3 ? (g.ntaps > 0 ? g.ntaps : 9) : 1;
  MTG_REQUIRE(ntaps >= 1 && ntaps <= 9, MTG_ERR_ARG, "conv_gemm: ntaps %d out of range", ntaps);
  for (int t = 0; t < 9; ++t) {
    kp.tap_dy[t] = static_cast<signed char>(g.ntaps > 0 ? g.tap_dy[t] : t / 3 - 1);
    kp.tap_dx[t] = static_cast<signed char>(g.ntaps > 0 ? g.tap_dx[t] : t % 3 - 1);
  }
  kp.num_kb = kp.kb_per_tap * ntaps;
  kp.ksteps_last = ceil_div(g.K - (kp.kb_per_tap - 1) * BK, 16);
  kp.scale = g.scale; kp.shift = g.shift; kp.act = g.act;
  kp.residual = g.residual; kp.out = g.out; kp.a_scale = g.a_scale; kp.hw = g.hw;
  kp.stat = g.stat;
  int tmem = 32;
  while (tmem < 2 * kp.BN) tmem <<= 1;
  kp.tmem_cols = tmem;

  kp.obox = (g.conv3x3 || kp.BN > 32) ? 64 : (kp.BN > 16 ? 32 : 16);  // narrow layers stage narrow slabs: more CTAs per SM
  kp.ss_bytes = SS_BYTES + (g.stat ? 2 * (512 / kp.obox) * kp.BN * 4 : 0);  // + [2][row groups][BN] partial sums
  kp.res_slabs = g.residual ? ceil_div(kp.BN, kp.obox) : 0;
  MTG_REQUIRE(!(g.conv3x3 && g.residual), MTG_ERR_UNSUPPORTED, "conv_gemm: residual with conv3x3 is not supported");
  // B-stationary: with few k-blocks the weights of one N tile fit in shared memory next to the A ring; every CTA then
  // streams only activations (the weight tile would otherwise be re-fetched from L2 for every 128-pixel tile)
  kp.m_tiles = g.conv3x3 ? 0 : ceil_div(g.M, BM);
  const size_t b_tile_bytes = static_cast<size_t>(kp.BN) * BK * 2;
  kp.b_res = (!g.conv3x3 && kp.num_kb <= 4 && kp.num_kb * b_tile_bytes <= 100 * 1024 &&
              static_cast<long long>(kp.m_tiles) * kp.n_tiles >= 4LL * num_sms()) ? 1 : 0;
  // per-warp epilogue (mode 1) needs the folded-BN constants of ONE N tile per CTA: a single N tile or resident weights
  static const bool legacy_epi = getenv("MTGSEG_GEMM_EPI") && atoi(getenv("MTGSEG_GEMM_EPI")) == 0;  // A/B switch
  // ... and enough tiles per CTA to keep both warp sets busy: with ~4 tiles per CTA (the 600-tile layers at 20x15) halving the
  // warps per tile costs more than the missing barriers save (measured: b9.expand 23 -> 34 us), from ~8 tiles on it wins
  // (b2.expand 198 -> 167 us, b3.expand 79 -> 71 us)
  const bool many_tiles = static_cast<long long>(kp.m_tiles) * kp.n_tiles >= 8LL * num_sms();
  kp.epi_mode = (!g.conv3x3 && !legacy_epi && many_tiles && (kp.n_tiles == 1 || kp.b_res)) ? 1 : 0;
  CUtensorMap tmA, tmB, tmO, tmR;
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
    const uint32_t box[4] = {(uint32_t)BK, (uint32_t)g.W, (uint32_t)kp.HB, (uint32_t)kp.NB};
    int rc = make_map(&tmA, g.a, 4, dims, strides, box);
    if (rc) return rc;
    const uint64_t wd[2] = {(uint64_t)g.K * ntaps, (uint64_t)g.N};
    const uint64_t ws[1] = {(uint64_t)g.K * ntaps * 2};
    const uint32_t wb[2] = {(uint32_t)BK, (uint32_t)kp.BN};
    rc = make_map(&tmB, g.w, 2, wd, ws, wb);
    if (rc) return rc;
    // output view: dense NHWC by default, or a strided view (transposed-conv parity classes write every other pixel)
    const uint64_t sx = g.out_sx ? g.out_sx : g.N, sy = g.out_sy ? g.out_sy : (uint64_t)g.W * g.N,
                   sn = g.out_sn ? g.out_sn : (uint64_t)g.H * g.W * g.N;
    const uint64_t od[4] = {(uint64_t)g.N, (uint64_t)g.W, (uint64_t)g.H, (uint64_t)g.B};
    const uint64_t os[3] = {sx * 2, sy * 2, sn * 2};
    const uint32_t ob[4] = {64, (uint32_t)g.W, (uint32_t)kp.HB, (uint32_t)kp.NB};
    rc = make_map(&tmO, g.out, 4, od, os, ob);
    if (rc) return rc;
    tmR = tmO;
  } else {
    kp.m_tiles = ceil_div(g.M, BM);
    kp.a_bytes = A_STAGE_BYTES;
    const uint64_t dims[2] = {(uint64_t)g.K, (uint64_t)g.M};
    const uint64_t strides[1] = {(uint64_t)g.K * 2};
    const uint32_t box[2] = {(uint32_t)BK, BM};
    int rc = make_map(&tmA, g.a, 2, dims, strides, box, BK);
    if (rc) return rc;
    const uint64_t wd[2] = {(uint64_t)g.K, (uint64_t)g.N};
    const uint64_t ws[1] = {(uint64_t)g.K * 2};
    const uint32_t wb[2] = {(uint32_t)BK, (uint32_t)kp.BN};
    rc = make_map(&tmB, g.w, 2, wd, ws, wb, BK);
    if (rc) return rc;
    const uint64_t od[2] = {(uint64_t)g.N, (uint64_t)g.M};
    const uint64_t os[1] = {(uint64_t)g.N * 2};
    const uint32_t ob[2] = {(uint32_t)kp.obox, BM};
    const uint32_t ob_warp[2] = {(uint32_t)kp.obox, 32};  // mode 1: every epilogue warp stores its own 32 rows
    rc = make_map(&tmO, g.out, 2, od, os, kp.epi_mode ? ob_warp : ob, kp.obox);
    if (rc) return rc;
    tmR = tmO;
    if (g.residual) {
      rc = make_map(&tmR, g.residual, 2, od, os, ob, kp.obox);
      if (rc) return rc;
    }
  }

  const size_t b_res_bytes = kp.b_res ? align_up(kp.num_kb * b_tile_bytes, 1024) : 0;
  const int stage_bytes = A_STAGE_BYTES + (kp.b_res ? 0 : kp.BN * BK * 2);
  const size_t slab_bytes = static_cast<size_t>(BM) * kp.obox * 2;
  // a long mainloop (>= 6 k-blocks) hides the reload of a single residual buffer; the 48-96 KB saved become ring stages
  kp.res_bufs = (kp.res_slabs && kp.num_kb >= 6) ? 1 : 2;
  // small residual tiles (narrow layers) in the per-warp epilogue mode: the residual stream needs as much prefetch depth as the
  // A ring has, or every tile waits a full HBM round trip for its residual (b1.project: 3.3 TB/s with two buffers); up to
  // 32 KB of ring, an even number of slots
  if (kp.res_slabs && kp.epi_mode && kp.res_bufs == 2) {
    int rbufs = static_cast<int>((32 * 1024) / (kp.res_slabs * slab_bytes)) & ~1;
    if (rbufs > MAX_RES) rbufs = MAX_RES;
    if (rbufs > 2) kp.res_bufs = rbufs;
  }
  // ring depth for a given amount of output staging: enough stages that two co-resident CTAs keep >= ~64 KB of loads in
  // flight per SM (HBM latency x bandwidth); prefer two co-resident CTAs over a deeper ring when that is what it costs
  auto plan = [&](size_t out_bytes, int& stages_out, size_t& need_out) {
    const size_t fixed = 2048 /*two 1024-byte alignments*/ + out_bytes + static_cast<size_t>(kp.res_bufs) * kp.res_slabs * slab_bytes + kp.ss_bytes + BAR_BYTES + b_res_bytes;
    int stages = kp.num_kb >= 8 ? 6 : 4;
    if (stage_bytes <= 16 * 1024) stages = 8;
    while (stages > 2 && stages * static_cast<size_t>(stage_bytes) + fixed > 227 * 1024) --stages;
    if (stages > 3 && 2 * (3 * static_cast<size_t>(stage_bytes) + fixed) <= 227 * 1024 &&
        2 * (stages * static_cast<size_t>(stage_bytes) + fixed) > 227 * 1024)
      stages = 3;
    stages_out = stages;
    need_out = static_cast<size_t>(stages) * stage_bytes + fixed;
  };
  int stages = 0;
  size_t need = 0;
  if (kp.epi_mode) {
    // 8 warps x 32 rows = two slabs' worth of staging per buffer; take the second buffer per warp only when it costs
    // neither a co-resident CTA nor ring stages below 4
    int st1, st2;
    size_t need1, need2;
    plan(2 * slab_bytes, st1, need1);
    plan(4 * slab_bytes, st2, need2);
    const bool two_ok = need2 <= 227 * 1024 && ((227 * 1024) / (need2 + 1024) >= 2) == ((227 * 1024) / (need1 + 1024) >= 2) && (st2 >= 4 || st2 == st1);
    kp.wbufs = two_ok ? 2 : 1;
    kp.out_bytes = static_cast<int>((two_ok ? 4 : 2) * slab_bytes);
    stages = two_ok ? st2 : st1;
    need = two_ok ? need2 : need1;
  } else {
    kp.wbufs = 0;
    kp.out_bytes = static_cast<int>(OUT_BUFS * slab_bytes);
    plan(OUT_BUFS * slab_bytes, stages, need);
  }
  kp.stages = stages;
  MTG_REQUIRE(need <= 227 * 1024, MTG_ERR_UNSUPPORTED, "conv_gemm: tile needs %zu B shared memory", need);
  if (g.conv3x3) return launch_variant<true, false>(tmA, tmB, tmO, tmR, kp, need, st);
  if (g.a_scale) return launch_variant<false, true>(tmA, tmB, tmO, tmR, kp, need, st);
  return launch_variant<false, false>(tmA, tmB, tmO, tmR, kp, need, st);
}

}  // namespace mtgseg
