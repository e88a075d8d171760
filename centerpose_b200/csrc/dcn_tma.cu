// TMA-staged DCNv2 (modulated deformable 3x3 convolution, stride 1, pad 1, dilation 1, one deformable group)
// on tcgen05, kind::tf32.  Reference semantics: dcn_v2_im2col_cuda.cu:125-195 (bilinear sampling with per-corner
// bounds, modulation mask) followed by the GEMM of dcn_v2_cuda.cu:105-160, + folded BatchNorm + ReLU
// (pose_dla_dcn.py:363-379 DeformConv).
//
// Why a second deformable kernel: the gather kernel (igemm_umma.cu) fetches the four bilinear corners of every
// (position, tap, channel chunk) from global memory; each input pixel is re-fetched ~36 times through L1/L2 and the
// producers sit on load latency (measured 1.08 ms for 64->64 @128x128, B=32: 16 B/clk/SM of L1 traffic).  Learned
// offsets are small, so here the rows a tile can reach are STAGED ONCE in shared memory by TMA and the corners are
// gathered from shared memory:
//
//   tile    = an 8 x 16 patch of output positions of one image (128 positions = the 128 TMEM lanes)
//   slab    = 16 channels x 32 columns x 24 rows around it (halo 8 in x and y, zero-filled outside the image),
//             64 bytes per position in TMA SWIZZLE_64B order (conflict-free 16-byte gathers); double buffered
//   K order = slab-major, tap-minor (the weight tiles of conv_tma.cu with 16-channel slabs are reused unchanged)
//   samples with a corner outside the slab (|offset| > ~7 px) are fetched from global memory instead - slow path,
//   same arithmetic, so the result never depends on how large the offsets are.
//
//   warp 0      : TMA producer for slabs
//   warp 1      : weight-tile producer (cp.async.bulk, pre-swizzled tiles)
//   warp 2      : tcgen05.mma issuer + TMEM owner
//   warp 3      : idle
//   warps 4-7   : epilogue (TMEM lane == position); x3: promotes every accumulation group into fp32 registers.
//                 Thread i also writes the sampling records (one 16-byte record per tap) of position i of the NEXT tile
//   warps 8-15  : gather, one thread per position: 4 corners x 16 channels from the slab -> bilinear blend * mask ->
//                 (hi, lo) tf32 split -> A tile in the UMMA SWIZZLE_64B K-major layout.  NG = 2 groups of four warps
//                 take K blocks kb = g, g + NG, ... and own AH A stages each.  (Measured, profiles/r02_dcn_ncu.md: every
//                 role of this pipeline waits 25-40 % of its time on its neighbour and the SM issues 43 % of its slots;
//                 16 gather warps in 4 groups at 80 registers were 15 % SLOWER -- the gather is not the limiter, the
//                 hand-offs are -- so the shared memory goes into a second A stage per group instead.)
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"
#include "umma_common.cuh"

namespace cp {
namespace {

using namespace umma;

constexpr int DT_BM = 128;
constexpr int DT_THREADS = 512;     // 4 control + 4 epilogue + 8 gather warps
constexpr int DT_CS = 16;            // channels per slab = one UMMA K block of 64-byte rows
constexpr int DT_PH = 8, DT_PW = 16;  // output patch (rows x columns) = 128 positions
constexpr int DT_HALO = 8;           // slab margin around the patch, both directions
constexpr int DT_SH = DT_PH + 2 * DT_HALO, DT_SW = DT_PW + 2 * DT_HALO;     // slab rows x columns (24 x 32)
constexpr uint32_t DT_SLAB_BYTES = DT_SH * DT_SW * DT_CS * 4;               // 49152
constexpr uint32_t DT_COEF_BYTES = DT_BM * 9 * 16;

struct DcnTmaParams {
  CUtensorMap amap;
  const float* src;
  int srcStride;
  const float* offmask;
  int omStride, mask_is_logit;
  int B, H, W, Cin, Cout, CoutPad, BN;
  int tiles_x, tiles_per_image;      // patches per image row / per image
  long long total_tiles;             // m tiles x n tiles (n fastest)
  int SB;
  int group;                         // x3: K blocks (16 channels) per TMEM accumulation group
  int AH;                            // A stages per gather group (1 or 2)
  int NG;                            // gather groups of four warps (2 or 4)
  int nbuf;                          // TMEM accumulation buffers (x3: one per in-flight accumulation group)
  const float* bias;
  const float* residual;
  int resStride, relu, res_after_relu;
  float* out;
  int outStride, out_nchw, round_tf32;
  const unsigned char* wtiles;
  // split-K (small maps: fewer tiles than SMs): tile = (m, n) tile * ksplit + split; a split walks `sps` slabs, stores its
  // partial sums to `part` ([mn tile][split][BN / 4][128 rows] float4) and dcn_tma_splitk_finish runs the epilogue
  int ksplit, sps;
  float* part;
};

struct DcnCtl {
  unsigned long long s_full[2], s_empty[2];
  unsigned long long a_full[8], a_empty[8];       // stages g AH .. g AH + AH - 1 belong to gather group g
  unsigned long long c_full[2], c_empty[2];
  unsigned long long b_full[8], b_empty[8];
  unsigned long long p_full[8], p_empty[8];
  uint32_t tmem_base;
};

// record bits: [0,14) slab position of the clamped top-left corner (in-slab) or (row << 7 | col) in the image
// (global path);  14: right corner is +1 column;  15: bottom corner is +1 row;  16-19: corner weights alive;
// 20: all corners inside the slab;  21: sample inside the image
constexpr int RB_DX = 14, RB_DY = 15, RB_W = 16, RB_SLAB = 20, RB_LIVE = 21;

// Sampling records of one output position (all 9 taps) -> shared memory.  dcn_v2_im2col_cuda.cu:160-195: the sample
// (h_im, w_im) is used only when it lies in (-1, H) x (-1, W); every corner carries its own bounds test.
__device__ __forceinline__ void coef_row(const DcnTmaParams& p, const float* __restrict__ om, int oy, int ox, int ys,
                                         int xs, uint32_t dst) {
  float o[27];
#pragma unroll
  for (int j = 0; j < 27; ++j) o[j] = __ldg(om + j);
  const int H = p.H, W = p.W;
#pragma unroll
  for (int tap = 0; tap < 9; ++tap) {
    const int ky = tap / 3, kx = tap - ky * 3;
    float mm = o[18 + tap];
    if (p.mask_is_logit) mm = 1.0f / (1.0f + expf(-mm));
    const float h_im = (float)(oy - 1 + ky) + o[2 * tap], w_im = (float)(ox - 1 + kx) + o[2 * tap + 1];
    float lh = 0.f, lw = 0.f, mk = 0.f;
    uint32_t pk = 0;
    if (h_im > -1.f && w_im > -1.f && h_im < (float)H && w_im < (float)W) {
      const int h_low = (int)floorf(h_im), w_low = (int)floorf(w_im);
      lh = h_im - (float)h_low;
      lw = w_im - (float)w_low;
      const bool t_ok = h_low >= 0, b_ok = h_low + 1 <= H - 1, l_ok = w_low >= 0, r_ok = w_low + 1 <= W - 1;
      const int hl = t_ok ? h_low : 0, hb = b_ok ? h_low + 1 : H - 1;
      const int wl = l_ok ? w_low : 0, wr = r_ok ? w_low + 1 : W - 1;
      const bool in_slab = (hl >= ys) && (hb < ys + DT_SH) && (wl >= xs) && (wr < xs + DT_SW);
      pk = in_slab ? (uint32_t)((hl - ys) * DT_SW + (wl - xs)) : (uint32_t)((hl << 7) | wl);
      pk |= (uint32_t)(wr - wl) << RB_DX;
      pk |= (uint32_t)(hb - hl) << RB_DY;
      pk |= (uint32_t)((t_ok && l_ok) ? 1 : 0) << (RB_W + 0);
      pk |= (uint32_t)((t_ok && r_ok) ? 1 : 0) << (RB_W + 1);
      pk |= (uint32_t)((b_ok && l_ok) ? 1 : 0) << (RB_W + 2);
      pk |= (uint32_t)((b_ok && r_ok) ? 1 : 0) << (RB_W + 3);
      pk |= (uint32_t)(in_slab ? 1 : 0) << RB_SLAB;
      pk |= 1u << RB_LIVE;
      mk = mm;
    }
    st_shared_v4(dst + (uint32_t)tap * 16u, __float_as_uint(lh), __float_as_uint(lw), __float_as_uint(mk), pk);
  }
}

template <bool X3>
__global__ void __launch_bounds__(DT_THREADS, 1) dcn_tma_kernel(const __grid_constant__ DcnTmaParams p) {
  extern __shared__ __align__(1024) unsigned char smem[];
  DcnCtl* ctl = reinterpret_cast<DcnCtl*>(smem);
  if (threadIdx.x == 0) griddep_launch_dependents();      // PDL (common.cuh)
  const uint32_t sbase = smem_u32(smem);
  const uint32_t coef0 = sbase + 1024u;
  const uint32_t slabs0 = (coef0 + 2u * DT_COEF_BYTES + 1023u) & ~1023u;
  const uint32_t a_stage = X3 ? 16384u : 8192u;                // hi (+ lo) tile of 128 rows x 64 bytes
  const uint32_t atiles0 = slabs0 + 2u * DT_SLAB_BYTES;
  const uint32_t btile_bytes = (uint32_t)p.BN * 64u * (X3 ? 2u : 1u);
  const uint32_t btiles0 = atiles0 + (uint32_t)(p.NG * p.AH) * a_stage;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n_tiles = p.CoutPad / p.BN;
  const int nslab = p.sps;                           // slabs one tile walks (all of them unless split-K)
  const int KB = nslab * 9;
  const int KB_all = (p.Cin / DT_CS) * 9;
  const int KS_SPLIT = p.ksplit;
  const long long total_tiles = p.total_tiles;

  if (tid == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(smem_u32(&ctl->s_full[s]), 1);
      mbar_init(smem_u32(&ctl->s_empty[s]), 4 * p.NG);      // one arrival per active gather warp
      mbar_init(smem_u32(&ctl->c_full[s]), 4);              // one arrival per epilogue warp
      mbar_init(smem_u32(&ctl->c_empty[s]), 4 * p.NG);
    }
    for (int s = 0; s < p.nbuf; ++s) {
      mbar_init(smem_u32(&ctl->p_full[s]), 1);
      mbar_init(smem_u32(&ctl->p_empty[s]), 128);
    }
    for (int s = 0; s < p.NG * p.AH; ++s) {
      mbar_init(smem_u32(&ctl->a_full[s]), 4);              // written by the four gather warps of one group
      mbar_init(smem_u32(&ctl->a_empty[s]), 1);
    }
    for (int s = 0; s < p.SB; ++s) {
      mbar_init(smem_u32(&ctl->b_full[s]), 1);
      mbar_init(smem_u32(&ctl->b_empty[s]), 1);
    }
    fence_mbar_init();
  }
  uint32_t tmem_cols = 32;
  // nbuf accumulation buffers of BN columns; x3 with an N tile > 64 keeps the promoted sums in BN more columns
  while ((int)tmem_cols < p.BN * (p.nbuf + ((X3 && p.BN > 64) ? 1 : 0))) tmem_cols <<= 1;
  if (warp == 2) {
    tmem_alloc(smem_u32(&ctl->tmem_base), tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = ctl->tmem_base;
  // PDL: the weight producer (warp 1) reads per-plan constants and runs ahead; every other role waits for the previous
  // launch (the offset / mask convolution whose output the records are built from) before it touches global memory
  if (warp != 1) griddep_wait();

  // warpgroup 0 (control warps) hands registers to warpgroup 1 (epilogue: 64 running sums + a 32-column TMEM read)
  if (warp < 4) asm volatile("setmaxnreg.dec.sync.aligned.u32 72;");
  if (warp == 0) {
    // ===================== slabs via TMA =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (long long tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const long long m_tile = (tile / KS_SPLIT) / n_tiles;
        const int s0 = (int)(tile % KS_SPLIT) * nslab;
        const int img = (int)(m_tile / p.tiles_per_image);
        const int pt = (int)(m_tile - (long long)img * p.tiles_per_image);
        const int y0 = (pt / p.tiles_x) * DT_PH, x0 = (pt % p.tiles_x) * DT_PW;
        for (int s = 0; s < nslab; ++s) {
          mbar_wait(smem_u32(&ctl->s_empty[stage]), phase ^ 1u);
          const uint32_t bar = smem_u32(&ctl->s_full[stage]);
          mbar_arrive_expect_tx(bar, DT_SLAB_BYTES);
          tma_load_4d(slabs0 + (uint32_t)stage * DT_SLAB_BYTES, &p.amap, (s0 + s) * DT_CS, x0 - DT_HALO, y0 - DT_HALO, img,
                      bar);
          if (++stage == 2) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================== weight tiles =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (long long tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int n_tile = (int)((tile / KS_SPLIT) % n_tiles);
        const unsigned char* wsrc = p.wtiles + ((size_t)n_tile * KB_all + (size_t)(tile % KS_SPLIT) * KB) * btile_bytes;
        for (int kb = 0; kb < KB; ++kb) {
          mbar_wait(smem_u32(&ctl->b_empty[stage]), phase ^ 1u);
          const uint32_t bar = smem_u32(&ctl->b_full[stage]);
          mbar_arrive_expect_tx(bar, btile_bytes);
          bulk_g2s(btiles0 + (uint32_t)stage * btile_bytes, wsrc + (size_t)kb * btile_bytes, btile_bytes, bar);
          if (++stage == p.SB) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 2) {
    // ===================== MMA issuer (warp-uniform loop, one elected lane issues) =====================
    // This loop is the critical path of the kernel (ncu: the MMA warp is busy 60 % of the time with ~130 dependent scalar
    // instructions per K block): descriptors are 32-bit low words + one shared high word, every parameter lives in a
    // register, and the two gather groups are handled by an explicit even / odd step.
    const uint32_t idesc = make_idesc_tf32(p.BN);
    const uint64_t dtmpl = make_desc(0, 0, DT_CS);
    const uint32_t dhi = (uint32_t)(dtmpl >> 32);
    const uint32_t dlo0 = (uint32_t)dtmpl;
    const uint32_t a_lo_u = 8192u >> 4;
    const uint32_t b_lo_u = ((uint32_t)p.BN * 64u) >> 4;
    const uint32_t a0_u = dlo0 + (atiles0 >> 4), a_stage_u = a_stage >> 4;
    const uint32_t b0_u = dlo0 + (btiles0 >> 4), b_stage_u = btile_bytes >> 4;
    const uint32_t bar_a_full = smem_u32(&ctl->a_full[0]), bar_a_empty = smem_u32(&ctl->a_empty[0]);
    const uint32_t bar_b_full = smem_u32(&ctl->b_full[0]), bar_b_empty = smem_u32(&ctl->b_empty[0]);
    const uint32_t bar_p_full = smem_u32(&ctl->p_full[0]), bar_p_empty = smem_u32(&ctl->p_empty[0]);
    const int SB = p.SB, AH = p.AH, group = p.group, nbuf = p.nbuf;
    const uint32_t bn = (uint32_t)p.BN;
    int sb = 0, buf = 0;
    uint32_t pb = 0, pe = 0;
    uint32_t cnt0 = 0, cnt1 = 0;          // K blocks consumed from gather group 0 / 1 (K block kbi of a tile -> group kbi & 1)
    for (long long tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      int gk = 0;
      for (int kbi = 0; kbi < KB; ++kbi) {
        const int grp = kbi & 1;
        const bool first = X3 ? (gk == 0) : (kbi == 0);
        if (first) mbar_wait(bar_p_empty + 8u * (uint32_t)buf, ((pe >> buf) & 1u) ^ 1u);
        const uint32_t hcnt = grp ? cnt1 : cnt0;
        const uint32_t sa = (uint32_t)(grp * AH) + (hcnt & (uint32_t)(AH - 1));
        mbar_wait(bar_a_full + 8u * sa, (hcnt >> (AH - 1)) & 1u);
        mbar_wait(bar_b_full + 8u * (uint32_t)sb, pb);
        tc_fence_after();
        const uint32_t da = a0_u + sa * a_stage_u;
        const uint32_t db = b0_u + (uint32_t)sb * b_stage_u;
        const uint32_t d_tmem = tmem_base + (uint32_t)buf * bn;
        const bool last = X3 ? (gk == group - 1 || kbi == KB - 1) : (kbi == KB - 1);
        if (elect_one()) {
          if (X3) {        // cross terms first, hi x hi last (see conv_tma.cu)
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) {
              umma_tf32_lohi(d_tmem, da + a_lo_u + 2u * ks, db + 2u * ks, dhi, idesc, (first && ks == 0) ? 0u : 1u);
              umma_tf32_lohi(d_tmem, da + 2u * ks, db + b_lo_u + 2u * ks, dhi, idesc, 1u);
            }
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) umma_tf32_lohi(d_tmem, da + 2u * ks, db + 2u * ks, dhi, idesc, 1u);
          } else {
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) umma_tf32_lohi(d_tmem, da + 2u * ks, db + 2u * ks, dhi, idesc, (first && ks == 0) ? 0u : 1u);
          }
          umma_commit(bar_b_empty + 8u * (uint32_t)sb);
          umma_commit(bar_a_empty + 8u * sa);
          if (last) umma_commit(bar_p_full + 8u * (uint32_t)buf);
        }
        __syncwarp();
        if (++sb == SB) {
          sb = 0;
          pb ^= 1u;
        }
        cnt0 += (uint32_t)(grp ^ 1);
        cnt1 += (uint32_t)grp;
        if (last) {
          pe ^= 1u << buf;
          if (++buf == nbuf) buf = 0;
          gk = 0;
        } else {
          ++gk;
        }
      }
    }
  } else if (warp == 3) {
    // idle
  } else if (warp >= 8) {
    // ===================== gather: thread = one position of the tile; the two halves of the 8 warps take alternate K
    // blocks (kb = half, half + 2, ...) and each half owns one A stage, so the record of a (position, tap) is decoded
    // once for all 16 channels and a stage is synchronised once per 128-thread task =====================
    const int gt = tid - 256;
    const int half = gt >> 7;              // gather group 0 .. NG - 1
    const int row = gt & 127;
    const long long g_tiles = half < p.NG ? total_tiles : 0;      // groups past NG idle (shared memory had room for two)
    const uint32_t a_row = atiles0 + (uint32_t)(row >> 3) * 512u + (uint32_t)(row & 7) * 64u;
    const uint32_t asw = (uint32_t)(row >> 1) & 3u;
    int ss = 0, cb = 0;
    uint32_t ps = 0, pc = 0;
    uint32_t cnt = 0;                      // K blocks this half has produced; stage = 2 half + (cnt & 1)
    for (long long tile = blockIdx.x; tile < g_tiles; tile += gridDim.x) {
      const long long m_tile = (tile / KS_SPLIT) / n_tiles;
      const int s0 = (int)(tile % KS_SPLIT) * nslab;
      const int img = (int)(m_tile / p.tiles_per_image);
      mbar_wait(smem_u32(&ctl->c_full[cb]), pc);
      const uint32_t crow = coef0 + (uint32_t)cb * DT_COEF_BYTES + (uint32_t)row * 144u;
      const float* gimg = p.src + (size_t)img * p.H * p.W * p.srcStride;
      int cur = -1;
      uint32_t slab = 0;
      float4 rec_next = ld_shared_v4f(crow + (uint32_t)(half % 9) * 16u);      // tap of this group's first K block
      for (int kb = half; kb < KB; kb += p.NG) {
        const int s = kb / 9;
        if (s != cur) {
          if (cur >= 0) {                     // done with the previous slab
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&ctl->s_empty[ss]));
            if (++ss == 2) {
              ss = 0;
              ps ^= 1u;
            }
          }
          mbar_wait(smem_u32(&ctl->s_full[ss]), ps);
          slab = slabs0 + (uint32_t)ss * DT_SLAB_BYTES;
          cur = s;
        }
        const float4 rec = rec_next;
        {                                   // the record of this group's next K block: its latency hides behind this gather
          const int kn = kb + p.NG;
          const int tn = kn - (kn / 9) * 9;
          rec_next = ld_shared_v4f(crow + (uint32_t)(kn < KB ? tn : 0) * 16u);
        }
        const uint32_t pk = __float_as_uint(rec.w);
        float4 v[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) v[c] = make_float4(0.f, 0.f, 0.f, 0.f);
        if ((pk >> RB_LIVE) & 1u) {
          const float lh = rec.x, lw = rec.y, mk = rec.z;
          const float hh = 1.f - lh, hw = 1.f - lw;
          const float w1 = ((pk >> (RB_W + 0)) & 1u) ? hh * hw * mk : 0.f;
          const float w2 = ((pk >> (RB_W + 1)) & 1u) ? hh * lw * mk : 0.f;
          const float w3 = ((pk >> (RB_W + 2)) & 1u) ? lh * hw * mk : 0.f;
          const float w4 = ((pk >> (RB_W + 3)) & 1u) ? lh * lw * mk : 0.f;
          if ((pk >> RB_SLAB) & 1u) {
            // slab rows are 64 bytes in TMA SWIZZLE_64B order: 16-byte chunk c of position q sits at c ^ ((q >> 1) & 3)
            const uint32_t q1 = pk & 0x3FFFu, q2 = q1 + ((pk >> RB_DX) & 1u);
            const uint32_t q3 = q1 + ((pk >> RB_DY) & 1u) * (uint32_t)DT_SW, q4 = q3 + ((pk >> RB_DX) & 1u);
            // chunk c of position q: slab + 64 q + ((c ^ s) << 4) == (slab + 64 q + (s << 4)) ^ (c << 4)  (64-byte aligned rows)
            const uint32_t b1 = slab + q1 * 64u + ((q1 << 3) & 0x30u), b2 = slab + q2 * 64u + ((q2 << 3) & 0x30u);
            const uint32_t b3 = slab + q3 * 64u + ((q3 << 3) & 0x30u), b4 = slab + q4 * 64u + ((q4 << 3) & 0x30u);
            // all sixteen 16-byte loads are issued before the first blend (ld_shared_v4f is volatile: source order is kept),
            // so their latencies overlap instead of being paid one after the other by this warp
            float4 c1[4], c2[4], c3[4], c4[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              c1[c] = ld_shared_v4f(b1 ^ ((uint32_t)c << 4));
              c2[c] = ld_shared_v4f(b2 ^ ((uint32_t)c << 4));
            }
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              c3[c] = ld_shared_v4f(b3 ^ ((uint32_t)c << 4));
              c4[c] = ld_shared_v4f(b4 ^ ((uint32_t)c << 4));
            }
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              v[c].x = w1 * c1[c].x + w2 * c2[c].x + w3 * c3[c].x + w4 * c4[c].x;
              v[c].y = w1 * c1[c].y + w2 * c2[c].y + w3 * c3[c].y + w4 * c4[c].y;
              v[c].z = w1 * c1[c].z + w2 * c2[c].z + w3 * c3[c].z + w4 * c4[c].z;
              v[c].w = w1 * c1[c].w + w2 * c2[c].w + w3 * c3[c].w + w4 * c4[c].w;
            }
          } else {      // corner rows outside the staged slab: same arithmetic from global memory
            const int hl = (int)((pk >> 7) & 127u), wl = (int)(pk & 127u);
            const float* g = gimg + ((size_t)hl * p.W + wl) * p.srcStride + (s0 + s) * DT_CS;
            const size_t dx = (size_t)((pk >> RB_DX) & 1u) * p.srcStride;
            const size_t dy = (size_t)((pk >> RB_DY) & 1u) * p.W * p.srcStride;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              const float4 c1 = __ldg(reinterpret_cast<const float4*>(g) + c);
              const float4 c2 = __ldg(reinterpret_cast<const float4*>(g + dx) + c);
              const float4 c3 = __ldg(reinterpret_cast<const float4*>(g + dy) + c);
              const float4 c4 = __ldg(reinterpret_cast<const float4*>(g + dy + dx) + c);
              v[c].x = w1 * c1.x + w2 * c2.x + w3 * c3.x + w4 * c4.x;
              v[c].y = w1 * c1.y + w2 * c2.y + w3 * c3.y + w4 * c4.y;
              v[c].z = w1 * c1.z + w2 * c2.z + w3 * c3.z + w4 * c4.z;
              v[c].w = w1 * c1.w + w2 * c2.w + w3 * c3.w + w4 * c4.w;
            }
          }
        }
        const int sa = half * p.AH + (int)(cnt & (uint32_t)(p.AH - 1));          // barrier index == tile slot
        const uint32_t a_hi = a_row + (uint32_t)sa * a_stage + (asw << 4);       // chunk c -> a_hi ^ (c << 4)
        mbar_wait(smem_u32(&ctl->a_empty[sa]), ((cnt >> (p.AH - 1)) & 1u) ^ 1u);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const uint32_t off = (uint32_t)c << 4;
          float4 h;
          h.x = tf32_round(v[c].x);
          h.y = tf32_round(v[c].y);
          h.z = tf32_round(v[c].z);
          h.w = tf32_round(v[c].w);
          st_shared_v4f(a_hi ^ off, h.x, h.y, h.z, h.w);
          if (X3)
            st_shared_v4f((a_hi ^ off) + 8192u, tf32_round(v[c].x - h.x), tf32_round(v[c].y - h.y), tf32_round(v[c].z - h.z),
                          tf32_round(v[c].w - h.w));
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&ctl->a_full[sa]));
        ++cnt;
      }
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(smem_u32(&ctl->s_empty[ss]));       // last slab of the tile
        mbar_arrive(smem_u32(&ctl->c_empty[cb]));
      }
      if (++ss == 2) {
        ss = 0;
        ps ^= 1u;
      }
      if (++cb == 2) {
        cb = 0;
        pc ^= 1u;
      }
    }
  } else {
    // ===================== epilogue (warps 4-7): TMEM lane == position of the tile =====================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 176;");
    const int q = warp & 3;
    const int i = q * 32 + lane;
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
    EpiParams ep;
    ep.bias = p.bias;
    ep.residual = p.residual;
    ep.resStride = p.resStride;
    ep.relu = p.relu;
    ep.res_after_relu = p.res_after_relu;
    ep.round_tf32 = p.round_tf32;
    ep.out = p.out;
    ep.outStride = p.outStride;
    ep.out_nchw = p.out_nchw;
    ep.Cout = p.Cout;
    ep.CoutPad = p.CoutPad;
    ep.H = p.H;
    ep.W = p.W;
    int buf = 0;
    uint32_t pf = 0;
    // The epilogue threads are idle most of a tile, and thread i == position i: they also prepare the sampling records
    // of the NEXT tile (double-buffered), so the gather warps never wait for them.
    int cb = 0;
    uint32_t pc = 0;
    auto make_records = [&](long long t) {
      const long long mt = (t / KS_SPLIT) / n_tiles;
      const int im = (int)(mt / p.tiles_per_image);
      const int pt = (int)(mt - (long long)im * p.tiles_per_image);
      const int y0 = (pt / p.tiles_x) * DT_PH, x0 = (pt % p.tiles_x) * DT_PW;
      const int oy = y0 + (i >> 4), ox = x0 + (i & 15);
      mbar_wait(smem_u32(&ctl->c_empty[cb]), pc ^ 1u);
      coef_row(p, p.offmask + ((size_t)((size_t)im * p.H + oy) * p.W + ox) * p.omStride, oy, ox, y0 - DT_HALO, x0 - DT_HALO,
               coef0 + (uint32_t)cb * DT_COEF_BYTES + (uint32_t)i * 144u);
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&ctl->c_full[cb]));
      if (++cb == 2) {
        cb = 0;
        pc ^= 1u;
      }
    };
    if ((long long)blockIdx.x < total_tiles) make_records(blockIdx.x);
    for (long long tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      if (tile + gridDim.x < total_tiles) make_records(tile + gridDim.x);
      const int n_tile = (int)((tile / KS_SPLIT) % n_tiles);
      const long long m_tile = (tile / KS_SPLIT) / n_tiles;
      const int n = (int)(m_tile / p.tiles_per_image);
      const int pt = (int)(m_tile - (long long)n * p.tiles_per_image);
      const int oy = (pt / p.tiles_x) * DT_PH + (i >> 4), ox = (pt % p.tiles_x) * DT_PW + (i & 15);
      const int m = (n * p.H + oy) * p.W + ox;
      const int col_end = min(p.Cout, (n_tile + 1) * p.BN);
      // split-K: the 32 finished columns of this position go to the workspace instead of through the epilogue
      float4* part_row = reinterpret_cast<float4*>(p.part) + (size_t)tile * (p.BN >> 2) * 128 + i;
      auto emit = [&](float (&vv)[32], int c0) {
        if (KS_SPLIT > 1) {
#pragma unroll
          for (int j = 0; j < 8; ++j)
            if (c0 + 4 * j < p.BN)
              __stcg(part_row + (size_t)((c0 >> 2) + j) * 128, make_float4(vv[4 * j], vv[4 * j + 1], vv[4 * j + 2], vv[4 * j + 3]));
        } else {
          epilogue_sub_tile(ep, nullptr, vv, lane, true, m, n, oy, ox, n_tile * p.BN + c0, col_end);
        }
      };
      if (X3) {
        // two-level accumulation (see conv_tma.cu): every finished group of <= 36 MMAs is added, with round-to-nearest,
        // into running sums.  The sums live in a third TMEM region (columns [2 BN, 3 BN)) instead of registers, so the
        // tile can be 128 columns wide with 128 epilogue threads: ld group + ld sum -> add -> st sum, 16 columns at a time.
        const uint32_t sum_base = lane_base + (uint32_t)(p.nbuf * p.BN);
        const int ngroups = (KB + p.group - 1) / p.group;
        if (p.BN <= 64) {
          // N tile <= 64: the running sums fit in 64 registers -- one TMEM read per group instead of two reads + a write
          float sums[64];
#pragma unroll
          for (int j = 0; j < 64; ++j) sums[j] = 0.f;
          for (int gi = 0; gi < ngroups; ++gi) {
            mbar_wait(smem_u32(&ctl->p_full[buf]), (pf >> buf) & 1u);
            tc_fence_after();
#pragma unroll
            for (int c = 0; c < 2; ++c) {
              if (c * 32 < p.BN) {
                uint32_t rr[32];
                tmem_ld16(lane_base + (uint32_t)(buf * p.BN + c * 32), rr);
                if (c * 32 + 16 < p.BN) {
                  tmem_ld16(lane_base + (uint32_t)(buf * p.BN + c * 32 + 16), rr + 16);
                } else {
#pragma unroll
                  for (int j = 16; j < 32; ++j) rr[j] = 0u;
                }
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 32; ++j) sums[c * 32 + j] += __uint_as_float(rr[j]);
              }
            }
            tc_fence_before();
            mbar_arrive(smem_u32(&ctl->p_empty[buf]));
            pf ^= 1u << buf;
            if (++buf == p.nbuf) buf = 0;
          }
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            if (c * 32 < p.BN) {
              float vv[32];
#pragma unroll
              for (int j = 0; j < 32; ++j) vv[j] = sums[c * 32 + j];
              emit(vv, c * 32);
            }
          }
          continue;
        }
        for (int gi = 0; gi < ngroups; ++gi) {
          mbar_wait(smem_u32(&ctl->p_full[buf]), (pf >> buf) & 1u);
          tc_fence_after();
          for (int c = 0; c * 16 < p.BN; ++c) {
            uint32_t rr[16], ss[16];
            tmem_ld16(lane_base + (uint32_t)(buf * p.BN + c * 16), rr);
            if (gi > 0) tmem_ld16(sum_base + (uint32_t)(c * 16), ss);
            tmem_ld_wait();
            if (gi > 0) {
#pragma unroll
              for (int j = 0; j < 16; ++j) rr[j] = __float_as_uint(__uint_as_float(ss[j]) + __uint_as_float(rr[j]));
            }
            tmem_st16(sum_base + (uint32_t)(c * 16), rr);
          }
          tmem_st_wait();
          tc_fence_before();
          mbar_arrive(smem_u32(&ctl->p_empty[buf]));
          pf ^= 1u << buf;
          if (++buf == p.nbuf) buf = 0;
        }
        for (int c0 = 0; c0 < p.BN; c0 += 32) {
          uint32_t rr[32];
          tmem_ld16(sum_base + (uint32_t)c0, rr);
          if (c0 + 16 < p.BN) {
            tmem_ld16(sum_base + (uint32_t)(c0 + 16), rr + 16);
          } else {
#pragma unroll
            for (int j = 16; j < 32; ++j) rr[j] = 0u;
          }
          tmem_ld_wait();
          float vv[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) vv[j] = __uint_as_float(rr[j]);
          emit(vv, c0);
        }
      } else {
        mbar_wait(smem_u32(&ctl->p_full[buf]), (pf >> buf) & 1u);
        tc_fence_after();
        for (int c0 = 0; c0 < p.BN; c0 += 32) {
          uint32_t rr[32];
          tmem_ld16(lane_base + (uint32_t)(buf * p.BN + c0), rr);
          if (c0 + 16 < p.BN) {
            tmem_ld16(lane_base + (uint32_t)(buf * p.BN + c0 + 16), rr + 16);
          } else {
#pragma unroll
            for (int j = 16; j < 32; ++j) rr[j] = 0u;
          }
          tmem_ld_wait();
          if (c0 + 32 >= p.BN) {
            tc_fence_before();
            mbar_arrive(smem_u32(&ctl->p_empty[buf]));
          }
          float vv[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) vv[j] = __uint_as_float(rr[j]);
          emit(vv, c0);
        }
        pf ^= 1u << buf;
        if (++buf == p.nbuf) buf = 0;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, tmem_cols);
}

// split-K, second half: one thread per (position, 4 channels) adds the partial sums in split order + epilogue
__global__ void __launch_bounds__(256) dcn_tma_splitk_finish(const __grid_constant__ DcnTmaParams p, long long mn_tiles) {
  griddep_launch_dependents();
  griddep_wait();
  const int G = p.BN >> 2;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= mn_tiles * G * 128) return;
  const int i = (int)(idx & 127);
  const int c4 = (int)((idx >> 7) % G);
  const long long mn = idx / (128ll * G);
  const float4* src = reinterpret_cast<const float4*>(p.part) + ((size_t)mn * p.ksplit * G + c4) * 128 + i;
  float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int q = 0; q < p.ksplit; ++q) {
    const float4 v = __ldcg(src + (size_t)q * G * 128);
    a.x += v.x;
    a.y += v.y;
    a.z += v.z;
    a.w += v.w;
  }
  const int n_tiles = p.CoutPad / p.BN;
  const int n_tile = (int)(mn % n_tiles);
  const long long m_tile = mn / n_tiles;
  const int n = (int)(m_tile / p.tiles_per_image);
  const int pt = (int)(m_tile - (long long)n * p.tiles_per_image);
  const int oy = (pt / p.tiles_x) * DT_PH + (i >> 4), ox = (pt % p.tiles_x) * DT_PW + (i & 15);
  const size_t m = ((size_t)n * p.H + oy) * p.W + ox;
  const int col0 = n_tile * p.BN + c4 * 4;
  const int col_end = min(p.Cout, (n_tile + 1) * p.BN);
  float v[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const bool in = col0 + j < col_end;
    if (col0 + j < p.CoutPad) v[j] += __ldg(p.bias + col0 + j);
    if (p.residual && in && !p.res_after_relu) v[j] += __ldg(p.residual + m * p.resStride + col0 + j);
    if (p.relu) v[j] = fmaxf(v[j], 0.f);
    if (p.residual && in && p.res_after_relu) v[j] += __ldg(p.residual + m * p.resStride + col0 + j);
    if (p.round_tf32) v[j] = tf32_round(v[j]);
  }
  if (p.out_nchw) {
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (col0 + j < col_end) p.out[(((size_t)n * p.Cout + col0 + j) * p.H + oy) * p.W + ox] = v[j];
  } else {
    float* o = p.out + m * p.outStride + col0;
    if (col0 + 3 < col_end) {
      *reinterpret_cast<float4*>(o) = make_float4(v[0], v[1], v[2], v[3]);
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (col0 + j < col_end) o[j] = v[j];
    }
  }
}

}  // namespace

int dcn_tma_tile_n(int CoutPad, int x3) {
  const int cap = x3 ? 128 : 256;     // x3: 2 BN columns of group accumulators + BN columns of promoted sums <= 512
  return CoutPad <= cap ? CoutPad : cap;
}

bool dcn_tma_supported(const IgemmParams& p, int x3) {
  if (p.mode != IGEMM_DCN || p.nsrc != 1) return false;
  if (p.kh != 3 || p.kw != 3 || p.stride != 1 || p.pad != 1) return false;
  if (p.Cin % DT_CS || p.srcStride[0] % 4) return false;
  const int W = p.Win, H = p.Hin;
  if (W > 128 || H > 128 || (W % DT_PW) || (H % DT_PH)) return false;     // records pack image coordinates in 7 bits
  const int bn = dcn_tma_tile_n(p.CoutPad, x3);
  if (bn % 16 || p.CoutPad % bn) return false;
  return true;
}

int dcn_tma_encode(const IgemmParams& p, int Bmax, void* map_out) {
  return tma_encode_nhwc_box(p.src[0], p.srcC[0], p.Win, p.Hin, Bmax, p.srcStride[0], DT_CS, DT_SW, DT_SH, 1, map_out);
}

int launch_dcn_tma(const IgemmParams& p, const void* map, int x3, int round_out_tf32, cudaStream_t stream) {
  if (!p.wgt_umma) return fail(CP_ERR_INVALID, "dcn_tma: weight tiles missing");
  if (!dcn_tma_supported(p, x3)) return fail(CP_ERR_INVALID, "dcn_tma: unsupported shape");
  DcnTmaParams q;
  memset(&q, 0, sizeof(q));
  memcpy(&q.amap, map, sizeof(CUtensorMap));
  q.src = p.src[0];
  q.srcStride = p.srcStride[0];
  q.offmask = p.offmask;
  q.omStride = p.omStride;
  q.mask_is_logit = p.mask_is_logit;
  q.B = p.B;
  q.H = p.Hin;
  q.W = p.Win;
  q.Cin = p.Cin;
  q.Cout = p.Cout;
  q.CoutPad = p.CoutPad;
  q.BN = dcn_tma_tile_n(p.CoutPad, x3);
  q.tiles_x = p.Win / DT_PW;
  q.tiles_per_image = q.tiles_x * (p.Hin / DT_PH);
  q.total_tiles = (long long)q.tiles_per_image * p.B * (p.CoutPad / q.BN);
  const uint32_t a_stage = x3 ? 16384u : 8192u;
  q.group = x3_group_blocks() * 2;      // 16-channel K blocks: same MMA count per group as conv_tma.cu
  const uint32_t btile = (uint32_t)q.BN * 64u * (x3 ? 2u : 1u);
  const size_t budget = 226 * 1024;
  // two A stages per gather group where >= 3 weight stages still fit (x3: N tile <= 64), else one
  q.NG = 2;
  q.AH = 2;
  size_t fixed = 1024 + 2 * (size_t)DT_COEF_BYTES + 1024 + 2 * (size_t)DT_SLAB_BYTES + (size_t)q.NG * q.AH * a_stage;
  if (getenv("CP_DCN_AH1") || fixed + 3 * (size_t)btile > budget) {
    q.AH = 1;
    fixed = 1024 + 2 * (size_t)DT_COEF_BYTES + 1024 + 2 * (size_t)DT_SLAB_BYTES + (size_t)q.NG * q.AH * a_stage;
  }
  if (fixed + 2 * (size_t)btile > budget) return fail(CP_ERR_INVALID, "dcn_tma: tile does not fit shared memory");
  // x3 hands a TMEM buffer to the promoting warps every 12 MMAs (~400 clk of tensor work) and each hand-off costs a
  // commit -> mbarrier -> tcgen05.ld -> arrive round trip of about the same length: with two buffers the MMA warp
  // spent 32 % of its time waiting for a free one (ncu, profiles/r02_dcn_ncu.md).  Use all 512 TMEM columns.
  q.nbuf = 2;
  if (x3) q.nbuf = q.BN <= 64 ? 8 : (512 - q.BN) / q.BN;
  if (const char* e = getenv("CP_DCN_NBUF")) q.nbuf = atoi(e) >= 2 && atoi(e) <= q.nbuf ? atoi(e) : q.nbuf;
  q.SB = (int)((budget - fixed) / btile);
  if (q.SB > 8) q.SB = 8;
  q.bias = p.bias;
  q.residual = p.residual;
  q.resStride = p.resStride;
  q.relu = p.relu;
  q.res_after_relu = p.res_after_relu;
  q.out = p.out;
  q.outStride = p.outStride;
  q.out_nchw = p.out_nchw;
  q.round_tf32 = round_out_tf32;
  q.wtiles = (const unsigned char*)p.wgt_umma;
  const size_t smem = fixed + (size_t)q.SB * btile;
  static PerDevice<bool, 2> configured;
  if (!configured.here(x3 ? 1 : 0)) {
    if (x3)
      CP_CUDA_CHECK(cudaFuncSetAttribute(dcn_tma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    else
      CP_CUDA_CHECK(cudaFuncSetAttribute(dcn_tma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    configured.here(x3 ? 1 : 0) = true;
  }
  int num_sms = 0;
  if (int rc = device_sm_count(&num_sms)) return rc;
  // split-K: at batch 1 the 512 -> 256 DCN at 16 x 16 is 4 tiles of 288 K blocks; deal slab ranges to idle SMs
  q.ksplit = 1;
  q.sps = p.Cin / DT_CS;
  q.part = p.splitk_ws;
  const long long mn = q.total_tiles;
  const char* ks_off = getenv("CP_NO_SPLITK");        // "1": no split-K anywhere, "dcn": not here, "conv": not in conv_tma
  if (p.splitk_ws && !(ks_off && (ks_off[0] == '1' || ks_off[0] == 'd'))) {
    const int nslab = p.Cin / DT_CS;
    int S = 1;
    for (int cand = 2; cand <= nslab; ++cand)
      if (nslab % cand == 0 && mn * cand <= num_sms && (size_t)mn * cand * 128 * q.BN <= p.splitk_ws_floats) S = cand;
    if (S > 1) {
      q.ksplit = S;
      q.sps = nslab / S;
      q.total_tiles = mn * S;
    }
  }
  const unsigned grid = (unsigned)(q.total_tiles < num_sms ? q.total_tiles : num_sms);
  if (x3)
    CP_CUDA_CHECK(launch_kernel(dcn_tma_kernel<true>, dim3(grid), dim3(DT_THREADS), smem, stream, q));
  else
    CP_CUDA_CHECK(launch_kernel(dcn_tma_kernel<false>, dim3(grid), dim3(DT_THREADS), smem, stream, q));
  CP_LAUNCH_CHECK("dcn_tma_kernel");
  if (q.ksplit > 1) {
    const long long threads = mn * (q.BN / 4) * 128;
    CP_CUDA_CHECK(launch_kernel(dcn_tma_splitk_finish, dim3((unsigned)((threads + 255) / 256)), dim3(256), 0, stream, q, mn));
    CP_LAUNCH_CHECK("dcn_tma_splitk_finish");
  }
  return CP_OK;
}

}  // namespace cp
