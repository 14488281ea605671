// Dense 3x3 convolution (stride 1, pad 1) as an implicit GEMM on the sm_100a tensor cores, haloed-tile variant.
//
//   out[n][y][x][:] = act( (sum over taps, k  in[n][y+dy][x+dx][k] * w[:][tap][k]) * scale + shift )
//
// What bounds the nine-shifted-boxes version in gemm_tc.cu is the L2 -> shared-memory fill, not the tensor pipe: every
// 128-pixel tile re-fetches the whole weight slab and the nine taps re-load the same pixels nine times (head 3x3 of
// train/model.py:110 at B = 256: 2.7 GB of ring traffic for 169 MB of operands).  Here
// * ONE haloed activation tile per 64-channel slab serves all nine taps.  The tile is a TMA box {64 ch, P = W + 1 columns
//   starting at x = -1, R + 2 rows starting at y0 - 1}: shared-memory row p = r * P + c.  Column 0 of every row is the
//   zero that TMA fills for x = -1 and doubles as the x = W padding of the row before it, so tap (dy, dx) of output
//   position m = r * P + x is simply row m + (dy + 1) * P + (dx + 1): the A operand of a tap is the SAME tile with the
//   descriptor start address advanced by that many 128-byte rows (the 128B swizzle is a function of the absolute
//   shared-memory address, which is also why advancing by 32 B along K works; measured: the descriptor's base-offset field
//   must stay 0 for these unaligned starts, filling it with (address >> 7) & 7 gives wrong results).  Output positions with c == W are junk and
//   are clipped by the TMA store (its box is {64, P, R}, x = W is outside the tensor).
// * every weight tile (one tap x 64 channels x BN outputs) is used for TWO M tiles (a "pair": 2 x 128 output positions,
//   two accumulators in TMEM), which halves the weight traffic per MMA.
// Activations and weights travel in two separate mbarrier rings (A: one stage per 64-channel slab, B: one per tap).
// Warp roles: w0 TMA producer, w1 MMA issuer + TMEM owner, w2..9 epilogue (TMEM -> folded BN / activation -> bf16 ->
// swizzled slab -> TMA store; optional per-channel BatchNorm statistics for the training forward).
#include <cuda.h>
#include <stdio.h>
#include <stdlib.h>

#include "ops.h"
#include "ptx.cuh"

namespace mtgseg {

int make_tma_map_bf16(CUtensorMap* map, const void* base, int rank, const unsigned long long* dims,
                      const unsigned long long* strides_bytes, const unsigned* box, int kbox);  // gemm_tc.cu

namespace {

constexpr int BM = 128;         // UMMA M = output positions (incl. junk columns) per tile
constexpr int TILES = 2;        // M tiles that share one weight stage
constexpr int MAX_SB = 8;
constexpr int MAX_SA = 3;
constexpr int SLAB_BYTES = BM * 128;  // 128 positions x 64 channels bf16
constexpr int THREADS = 320;

struct C3Params {
  int N, K, BN, n_tiles;
  int B, H, W, P, R, h_tiles, m_tiles, units;
  int kc_count, ksteps_last;
  int sa, sb;
  int a_tile_bytes;  // shared memory per haloed tile (1024-aligned, covers the furthest row any tap can address)
  int a_box_bytes;   // bytes one TMA box delivers: (R + 2) * P rows of 128 B
  int tmem_cols;
  uint32_t desc_hi;
  const float* scale;
  const float* shift;
  int act;
  double* stat;
};

__global__ void __launch_bounds__(THREADS, 1)
conv3x3_halo_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const __grid_constant__ CUtensorMap tmO, const C3Params p) {
  extern __shared__ uint8_t smem_raw[];
  pdl_trigger();
  uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  const int b_stage_bytes = p.BN * 128;
  uint8_t* sA = smem;                                          // [sa][TILES][a_tile_bytes]
  uint8_t* sB = sA + p.sa * TILES * p.a_tile_bytes;            // [sb][BN rows x 128 B]
  uint8_t* sOut = sB + ((p.sb * b_stage_bytes + 1023) & ~1023);  // two staging slabs
  float* sScale = reinterpret_cast<float*>(sOut + 2 * SLAB_BYTES);
  float* sShift = sScale + 128;
  float* sSum = sShift + 128;        // [8 row groups][BN] (only with p.stat)
  float* sSq = sSum + 8 * 128;
  uint8_t* sValid = reinterpret_cast<uint8_t*>(sSq + 8 * 128);  // [2][128]: does output position m of the current tile exist
  uint64_t* bars = reinterpret_cast<uint64_t*>(sValid + 256);
  uint64_t* a_full = bars;
  uint64_t* a_empty = a_full + MAX_SA;
  uint64_t* b_full = a_empty + MAX_SA;
  uint64_t* b_empty = b_full + MAX_SB;
  uint64_t* tfull = b_empty + MAX_SB;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.sa; ++s) { ptx::mbar_init(&a_full[s], 1); ptx::mbar_init(&a_empty[s], 1); }
    for (int s = 0; s < p.sb; ++s) { ptx::mbar_init(&b_full[s], 1); ptx::mbar_init(&b_empty[s], 1); }
    for (int b = 0; b < 2; ++b) { ptx::mbar_init(&tfull[b], 1); ptx::mbar_init(&tempty[b], 256); }
    ptx::fence_mbar_init();
  }
  if (warp == 0 && lane == 0) {
    ptx::tma_prefetch_desc(&tmA);
    ptx::tma_prefetch_desc(&tmB);
    ptx::tma_prefetch_desc(&tmO);
  }
  if (warp == 1) {
    ptx::tmem_alloc(tmem_slot, p.tmem_cols);
    ptx::tmem_relinquish();
  }
  // the rows behind the TMA box of every haloed tile are never written by TMA: row (R + 2) * P is the x = W padding of the
  // last halo row (a real operand), the rest only feed junk output positions.  Zero them once.
  {
    const int tail16 = (p.a_tile_bytes - p.a_box_bytes) >> 4;
    for (int t = 0; t < p.sa * TILES; ++t) {
      uint4* dst = reinterpret_cast<uint4*>(sA + t * p.a_tile_bytes + p.a_box_bytes);
      for (int i = threadIdx.x; i < tail16; i += THREADS) dst[i] = make_uint4(0, 0, 0, 0);
    }
    ptx::fence_proxy_async_smem();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  pdl_wait();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t buf_stride = TILES * p.BN;

  if (warp == 0) {
    // =============================== TMA producer ===============================
    if (lane == 0) {
      // the A ring runs up to one slab ahead of the B ring: the next slab's haloed tiles are requested as soon as their stage
      // is free, from inside the tap loop, so that neither ring ever waits for the other
      uint32_t a_it = 0;            // A stages issued so far
      int a_unit = blockIdx.x, a_kc = 0;
      auto issue_a = [&](bool block) -> bool {
        if (a_unit >= p.units) return true;
        const int s = a_it % p.sa;
        const uint32_t ph = (a_it / p.sa) & 1;
        if (block) ptx::mbar_wait(&a_empty[s], ph ^ 1);
        else if (!ptx::mbar_try_wait(&a_empty[s], ph ^ 1)) return false;
        const int pair = a_unit / p.n_tiles;
        const int nvalid = (2 * pair + 1 < p.m_tiles) ? 2 : 1;
        ptx::mbar_arrive_expect_tx(&a_full[s], nvalid * p.a_box_bytes);
        for (int t = 0; t < nvalid; ++t) {
          const int mt = 2 * pair + t;
          const int n = mt / p.h_tiles, y0 = (mt - n * p.h_tiles) * p.R;
          ptx::tma_load_4d(sA + (s * TILES + t) * p.a_tile_bytes, &tmA, &a_full[s], a_kc * 64, -1, y0 - 1, n);
        }
        ++a_it;
        if (++a_kc == p.kc_count) { a_kc = 0; a_unit += gridDim.x; }
        return true;
      };
      uint32_t b_it = 0, chunk = 0;  // chunk: (unit, kc) slabs whose taps have been started
      for (int unit = blockIdx.x; unit < p.units; unit += gridDim.x) {
        const int n_tile = unit % p.n_tiles;
        for (int kc = 0; kc < p.kc_count; ++kc, ++chunk) {
          while (a_it <= chunk) issue_a(true);
          for (int tap = 0; tap < 9; ++tap, ++b_it) {
            if (a_it == chunk + 1) issue_a(false);
            const int s = b_it % p.sb;
            const uint32_t ph = (b_it / p.sb) & 1;
            ptx::mbar_wait(&b_empty[s], ph ^ 1);
            ptx::mbar_arrive_expect_tx(&b_full[s], b_stage_bytes);
            ptx::tma_load_2d(sB + s * b_stage_bytes, &tmB, &b_full[s], tap * p.K + kc * 64, n_tile * p.BN);
          }
        }
      }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer ===============================
    // The whole warp walks the loops (warp-uniform control flow keeps stage indices and descriptors in uniform registers);
    // one elected lane issues the MMAs and commits.
    const uint32_t idesc = ptx::umma_idesc_bf16(BM, p.BN);
    const uint32_t sA_u32 = ptx::smem_u32(sA), sB_u32 = ptx::smem_u32(sB);
    uint32_t a_it = 0, b_it = 0, tc = 0;
    for (int unit = blockIdx.x; unit < p.units; unit += gridDim.x, ++tc) {
      const int pair = unit / p.n_tiles;
      const bool two = 2 * pair + 1 < p.m_tiles;
      const uint32_t buf = tc & 1, aph = (tc >> 1) & 1;
      ptx::mbar_wait(&tempty[buf], aph ^ 1);
      ptx::tc_fence_after();
      const uint32_t d = tmem_base + buf * buf_stride;
      for (int kc = 0; kc < p.kc_count; ++kc, ++a_it) {
        const int sa = a_it % p.sa;
        ptx::mbar_wait(&a_full[sa], (a_it / p.sa) & 1);
        const int ksteps = (kc == p.kc_count - 1) ? p.ksteps_last : 4;
        const uint32_t a0 = sA_u32 + sa * TILES * p.a_tile_bytes;
        for (int ty = 0; ty < 3; ++ty) {
          for (int tx = 0; tx < 3; ++tx, ++b_it) {
            const int sb = b_it % p.sb;
            ptx::mbar_wait(&b_full[sb], (b_it / p.sb) & 1);
            ptx::tc_fence_after();
            const uint32_t a_addr = a0 + (ty * p.P + tx) * 128;  // tap (dy, dx) starts (dy + 1) * P + (dx + 1) rows into the haloed tile
            const uint64_t adesc0 = ptx::umma_desc_kmajor(a_addr, p.desc_hi);
            const uint64_t adesc1 = ptx::umma_desc_kmajor(a_addr + p.a_tile_bytes, p.desc_hi);
            const uint64_t bdesc = ptx::umma_desc_kmajor(sB_u32 + sb * b_stage_bytes, p.desc_hi);
            if (ptx::elect_one()) {
#pragma unroll
              for (int k = 0; k < 4; ++k)
                if (k < ksteps) ptx::umma_bf16(d, adesc0 + 2 * k, bdesc + 2 * k, idesc, (kc | ty | tx | k) != 0);
              if (two) {
#pragma unroll
                for (int k = 0; k < 4; ++k)
                  if (k < ksteps) ptx::umma_bf16(d + p.BN, adesc1 + 2 * k, bdesc + 2 * k, idesc, (kc | ty | tx | k) != 0);
              }
              ptx::umma_commit(&b_empty[sb]);
            }
            __syncwarp();
          }
        }
        if (ptx::elect_one()) ptx::umma_commit(&a_empty[sa]);
        __syncwarp();
      }
      if (ptx::elect_one()) ptx::umma_commit(&tfull[buf]);
      __syncwarp();
    }
  } else {
    // =============================== epilogue (8 warps) ===============================
    const int q = warp & 3;              // TMEM lane quarter this warp may access
    const int half = (warp - 2) >> 2;    // which 32 of the 64 slab columns
    const int r = q * 32 + lane;         // accumulator row == output position inside the tile
    const int et = threadIdx.x - 64;     // 0..255
    const bool leader = et == 0;
    const int slabs = (p.BN + 63) >> 6;
    const int rx = r & 7;
    if (p.stat)
      for (int i = et; i < 2 * 8 * 128; i += 256) sSum[i] = 0.f;
    auto convert_half = [&](uint32_t taddr, int sl, int cols, uint8_t* srow) {
      uint32_t v[2][16];
      ptx::tmem_ld16(taddr + sl * 64 + half * 32, v[0]);
      if (half * 32 + 16 < cols) ptx::tmem_ld16(taddr + sl * 64 + half * 32 + 16, v[1]);
      ptx::tmem_ld_wait();
#pragma unroll
      for (int cc = 0; cc < 2; ++cc) {
        if (half * 32 + cc * 16 < cols) {
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int chunk = half * 4 + cc * 2 + h;
            const int c = sl * 64 + chunk * 8;
            float f[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) f[e] = fmaf(__uint_as_float(v[cc][h * 8 + e]), sScale[c + e], sShift[c + e]);
            if (p.act == ACT_RELU) {
#pragma unroll
              for (int e = 0; e < 8; ++e) f[e] = fmaxf(f[e], 0.f);
            } else if (p.act != ACT_NONE) {
#pragma unroll
              for (int e = 0; e < 8; ++e) f[e] = apply_act(f[e], p.act);
            }
            *reinterpret_cast<uint4*>(srow + ((chunk ^ rx) << 4)) = pack8(f);
          }
        }
      }
    };
    // per-channel sums of the bf16 values the TMA store writes, over the positions that exist (training BatchNorm statistics).
    // Thread et owns column pair et % 32 and rows [16 * (et / 32), +16): its own slot, plain adds, fixed order.
    auto slab_stats = [&](const uint8_t* sbuf, int sl, int cols, const uint8_t* valid) {
      const int cp = et & 31, grp = et >> 5;
      if (cp * 2 >= cols) return;
      float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
      const int chunk = cp >> 2, within = (cp & 3) * 4;
      for (int i = 0; i < 16; ++i) {
        const int row = grp * 16 + i;
        if (!valid[row]) continue;
        const uint32_t w = *reinterpret_cast<const uint32_t*>(sbuf + row * 128 + ((chunk ^ (row & 7)) << 4) + within);
        const float a = __uint_as_float(w << 16), b = __uint_as_float(w & 0xFFFF0000u);
        s0 += a; s1 += b;
        q0 = fmaf(a, a, q0); q1 = fmaf(b, b, q1);
      }
      const int slot = grp * 128 + sl * 64 + cp * 2;
      sSum[slot] += s0; sSum[slot + 1] += s1;
      sSq[slot] += q0; sSq[slot + 1] += q1;
    };
    auto flush_stats = [&](int n_tile) {
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (et < p.BN) {
        const int n = n_tile * p.BN + et;
        float a = 0.f, b = 0.f;
        for (int g = 0; g < 8; ++g) {
          a += sSum[g * 128 + et]; b += sSq[g * 128 + et];
          sSum[g * 128 + et] = 0.f; sSq[g * 128 + et] = 0.f;
        }
        if (n < p.N) {
          atomicAdd(p.stat + n, static_cast<double>(a));
          atomicAdd(p.stat + p.N + n, static_cast<double>(b));
        }
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
    };
    uint32_t tc = 0, store_no = 0, tile_no = 0;
    int cur_ntile = -1;
    for (int unit = blockIdx.x; unit < p.units; unit += gridDim.x, ++tc) {
      const int pair = unit / p.n_tiles, n_tile = unit - pair * p.n_tiles;
      const int nvalid = (2 * pair + 1 < p.m_tiles) ? 2 : 1;
      const uint32_t buf = tc & 1, aph = (tc >> 1) & 1;
      if (n_tile != cur_ntile) {  // folded-BN constants of this N tile (visible after the first bar.sync below)
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
      for (int t = 0; t < nvalid; ++t, ++tile_no) {
        const int mt = 2 * pair + t;
        const int n = mt / p.h_tiles, y0 = (mt - n * p.h_tiles) * p.R;
        uint8_t* valid = sValid + (tile_no & 1) * 128;
        if (p.stat && et < 128) {
          const int rr = et / p.P, cc = et - rr * p.P;
          valid[et] = (rr < p.R && cc < p.W && y0 + rr < p.H) ? 1 : 0;
        }
        const uint32_t taddr = tmem_base + buf * buf_stride + t * p.BN + (static_cast<uint32_t>(q * 32) << 16);
        for (int sl = 0; sl < slabs; ++sl, ++store_no) {
          uint8_t* sbuf = sOut + (store_no & 1) * SLAB_BYTES;
          if (leader) ptx::bulk_wait_read<1>();  // the store that used this slab two slabs ago has finished reading it
          asm volatile("bar.sync 1, 256;" ::: "memory");
          const int cols = min(64, p.BN - sl * 64);  // multiple of 16
          if (half * 32 < cols) convert_half(taddr, sl, cols, sbuf + r * 128);
          ptx::fence_proxy_async_smem();
          asm volatile("bar.sync 1, 256;" ::: "memory");
          if (leader) {
            ptx::tma_store_4d(&tmO, sbuf, n_tile * p.BN + sl * 64, 0, y0, n);
            ptx::bulk_commit();
          }
          if (p.stat) slab_stats(sbuf, sl, cols, valid);
        }
      }
      ptx::tc_fence_before();
      ptx::mbar_arrive(&tempty[buf]);
    }
    if (p.stat && cur_ntile >= 0) flush_stats(cur_ntile);
    if (leader) ptx::bulk_wait_all();
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

int env_int(const char* name, int dflt) {
  const char* v = getenv(name);
  return v ? atoi(v) : dflt;
}

}  // namespace

// Returns MTG_OK after launching, a negative error code on failure, or 1 when this shape is not handled here (the caller then
// falls back to the nine-shifted-boxes path of gemm_tc.cu).
int launch_conv3x3_halo(const ConvGemmArgs& g, cudaStream_t st) {
  static const int mode = env_int("MTGSEG_CONV3", 1);  // A/B: 0 = always the nine-shifted-boxes kernel
  if (!mode || !g.conv3x3 || g.ntaps != 0 || g.out_sx || g.out_sy || g.out_sn || g.residual || g.a_scale) return 1;
  if (g.W + 1 > 64 || g.K % 8 || g.N % 8) return 1;
  if (g.B <= 0 || g.H <= 0 || g.W <= 0 || static_cast<long long>(g.B) * g.H * g.W != g.M) return 1;  // the caller reports it
  C3Params kp{};
  kp.N = g.N; kp.K = g.K;
  kp.BN = g.N <= 128 ? static_cast<int>(align_up(g.N, 16)) : 128;
  kp.n_tiles = ceil_div(g.N, kp.BN);
  kp.B = g.B; kp.H = g.H; kp.W = g.W;
  kp.P = g.W + 1;
  int rmax = BM / kp.P;
  if (rmax > g.H) rmax = g.H;
  kp.h_tiles = ceil_div(g.H, rmax);
  kp.R = ceil_div(g.H, kp.h_tiles);  // fewest tiles, then the smallest halo box
  kp.m_tiles = kp.h_tiles * g.B;
  kp.units = ceil_div(kp.m_tiles, TILES) * kp.n_tiles;
  kp.kc_count = ceil_div(g.K, 64);
  kp.ksteps_last = ceil_div(g.K - (kp.kc_count - 1) * 64, 16);
  kp.a_box_bytes = (kp.R + 2) * kp.P * 128;
  kp.a_tile_bytes = static_cast<int>(align_up(static_cast<size_t>(BM + 2 * kp.P + 2) * 128, 1024));
  if (kp.a_tile_bytes < kp.a_box_bytes + 128) kp.a_tile_bytes = static_cast<int>(align_up(kp.a_box_bytes + 128, 1024));
  kp.desc_hi = (1024u >> 4) | (1u << 14) | (2u << 29);  // SBO = 1024 B, descriptor version 1, SWIZZLE_128B
  kp.scale = g.scale; kp.shift = g.shift; kp.act = g.act; kp.stat = g.stat;
  int tmem = 32;
  while (tmem < 2 * TILES * kp.BN) tmem <<= 1;
  kp.tmem_cols = tmem;
  const size_t b_stage = static_cast<size_t>(kp.BN) * 128;
  const size_t fixed = 1024 /*alignment*/ + 2 * SLAB_BYTES + 2 * 128 * 4 + 2 * 8 * 128 * 4 + 256 + 512 /*barriers*/ + 1024 /*B ring round-up*/;
  kp.sa = 2;
  const size_t a_ring = static_cast<size_t>(kp.sa) * TILES * kp.a_tile_bytes;
  if (fixed + a_ring + 3 * b_stage > 227 * 1024) return 1;
  kp.sb = static_cast<int>((227 * 1024 - fixed - a_ring) / b_stage);
  if (kp.sb > MAX_SB) kp.sb = MAX_SB;
  const size_t need = fixed + a_ring + kp.sb * b_stage;

  CUtensorMap tmA, tmB, tmO;
  {
    const unsigned long long dims[4] = {(unsigned long long)g.K, (unsigned long long)g.W, (unsigned long long)g.H, (unsigned long long)g.B};
    const unsigned long long strides[3] = {(unsigned long long)g.K * 2, (unsigned long long)g.W * g.K * 2, (unsigned long long)g.H * g.W * g.K * 2};
    const unsigned box[4] = {64u, (unsigned)kp.P, (unsigned)(kp.R + 2), 1u};
    int rc = make_tma_map_bf16(&tmA, g.a, 4, dims, strides, box, 64);
    if (rc) return rc;
    const unsigned long long wd[2] = {(unsigned long long)g.K * 9, (unsigned long long)g.N};
    const unsigned long long ws[1] = {(unsigned long long)g.K * 9 * 2};
    const unsigned wb[2] = {64u, (unsigned)kp.BN};
    rc = make_tma_map_bf16(&tmB, g.w, 2, wd, ws, wb, 64);
    if (rc) return rc;
    const unsigned long long od[4] = {(unsigned long long)g.N, (unsigned long long)g.W, (unsigned long long)g.H, (unsigned long long)g.B};
    const unsigned long long os[3] = {(unsigned long long)g.N * 2, (unsigned long long)g.W * g.N * 2, (unsigned long long)g.H * g.W * g.N * 2};
    const unsigned ob[4] = {64u, (unsigned)kp.P, (unsigned)kp.R, 1u};
    rc = make_tma_map_bf16(&tmO, g.out, 4, od, os, ob, 64);
    if (rc) return rc;
  }
  static bool configured = false;
  if (!configured) {
    MTG_CUDA(cudaFuncSetAttribute(conv3x3_halo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    configured = true;
  }
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
  }
  // one CTA per SM (512 TMEM columns at BN = 128): pad the request so that a second CTA can never become co-resident
  size_t smem = need;
  if (smem < 116 * 1024) smem = 116 * 1024;
  int grid = kp.units < sms ? kp.units : sms;
  static const bool debug = getenv("MTGSEG_GEMM_DEBUG") != nullptr;
  if (debug)
    fprintf(stderr, "[conv3x3_halo] B=%d %dx%d K=%d N=%d BN=%d P=%d R=%d h_tiles=%d m_tiles=%d units=%d sa=%d sb=%d a_tile=%d a_box=%d tmem=%d smem=%zu grid=%d\n",
            g.B, g.H, g.W, g.K, g.N, kp.BN, kp.P, kp.R, kp.h_tiles, kp.m_tiles, kp.units, kp.sa, kp.sb, kp.a_tile_bytes, kp.a_box_bytes,
            kp.tmem_cols, smem, grid);
  MTG_CUDA(launch_pdl(conv3x3_halo_kernel, dim3(grid), dim3(THREADS), smem, st, tmA, tmB, tmO, kp));
  MTG_LAUNCH_CHECK();
  return MTG_OK;
}

}  // namespace mtgseg
