// gf_conv.cu -- implicit-GEMM 3-D convolution for the Wan video VAE on sm_100a (tcgen05 / TMEM / TMA).
//
//   Y[to, ho, wo, :] = epilogue( sum_{dt,dh,dw,ci} Wt[co, (dt,dh,dw), ci] * X[to*st + dt - pt, ho*sh + dh - ph, wo*sw + dw - pw, ci] )
//
// Replaces CausalConv3d (diffsynth/models/wan_video_vae.py:33-52), the Conv2d of Resample (:92-119) and the strided
// time_conv (:104-119) of the reference.  The clip is channels-last ([T][H][W][C] bf16), so the GEMM is
//   M = To*Ho*Wo output positions,  N = Cout,  K = taps * Cin,
// and nothing is ever gathered into an im2col buffer: the TMA producer loads boxes of ONE 4-D tensor map (C, W, H, T) at
// shifted coordinates, and box elements outside the clip (negative t: the causal front padding; h, w outside: the
// spatial zero padding) are zero-filled by TMA.
//
// Two kernels (DESIGN 10 has the shared-memory-bandwidth model that decides between them):
//   gf_conv3d_halo_kernel  3x3 spatial windows, stride 1 (95 % of the VAE's FLOPs): one input halo per (dt, channel
//                          block) serves all nine taps through row-shifted UMMA descriptors; CTA pairs (cta_group::2).
//   gf_conv3d_kernel       everything else (strided, 1x1, (3,1,1), folded windows): one 128-row box per tap, a tensor
//                          map per input parity for stride 2; gf_gemm.cu's single-CTA skeleton (warp 0 TMA producer,
//                          warp 1 MMA issuer + TMEM owner, warps 2..5 epilogue, two 256-column accumulator stages).
// A k-block is 64 channels of one tap; for Cin % 64 != 0 the last block of a tap issues only the valid 16-wide MMAs
// (the TMA box is zero-filled, the weights beside it are never multiplied).
//
// Epilogue, shared by both (conv_epilogue_rows; per output position, fp32 accumulator):
//   v  = bf16(acc + bias)                      ; v = bf16(v + R) when a residual is given (ResidualBlock tail, :296-301)
//   Y  = v                                     (channels-last, or (C, T, H, W) planes for Cout <= 32)
//   Y2 = silu(v / max(|v|_2, 1e-12) * sqrt(C) * gamma)   (the NEXT layer's RMS_norm + SiLU, :55-70,283-288) when the
//        whole channel row lives in one tile (Cout <= 256); the row stays in registers between the two passes.
#include <type_traits>
#include "gf_ptx.cuh"
#include "gf_api_internal.h"

namespace gf {

constexpr int CONV_BM = 128;
constexpr int CONV_BK = 64;
constexpr int CONV_THREADS = 192;
constexpr int CONV_MAX_TAPS = 27;
constexpr int CONV_MAX_STAGES = 8;
constexpr int CONV_A_BYTES = CONV_BM * CONV_BK * 2;   // 16 KB
constexpr int CONV_SMEM_BUDGET = 200 * 1024;
constexpr int CONV_STG_PITCH = 80;                       // epilogue staging: bytes per row (32 bf16 + 16 B pad)
constexpr int CONV_STG_BYTES = 32 * CONV_STG_PITCH;      // per epilogue warp

struct ConvMaps {
  CUtensorMap a[4];   // input clip, one map per (h, w) parity of a strided convolution
  CUtensorMap b;      // weights [Cout][taps * Cin]
};

struct ConvParams {
  int To, Ho, Wo;
  int Cin;              // multiple of 8
  int Cout;             // real output channels
  int cout_store;       // Cout rounded up to 8 (channels-last rows are written in 16-byte vectors)
  int ntaps;
  int st;               // temporal stride
  int bw_shift;         // BW = 1 << bw_shift, BH = 128 >> bw_shift
  int nWt, nHt;
  int BN, num_n_tiles;
  int stages;
  int a_stages;         // halo form: input-halo stages
  int ncthw;            // 1: Y is (Cout, To, Ho, Wo)
  int silu;
  signed char tap_map[CONV_MAX_TAPS], tap_dt[CONV_MAX_TAPS], tap_dh[CONV_MAX_TAPS], tap_dw[CONV_MAX_TAPS];
  __nv_bfloat16* Y; long long ldy;
  const __nv_bfloat16* bias;
  const __nv_bfloat16* R; long long ldr;
  __nv_bfloat16* Y2; long long ldy2;
  const __nv_bfloat16* gamma;
};

// ------------------------------------------------------------------------------------------------------------------
// Epilogue of one warp for its 32 accumulator rows (TMEM lanes) and up to NCH x 32 columns starting at column n0.
// Pass 0 reads the accumulator, adds bias and residual (bf16 roundings of the torch ops), keeps the packed bf16 row in
// registers, stores Y and hands the accumulator back; pass 1 (fused RMS_norm + SiLU of the next layer) works from the
// registers.  R / Y / Y2 move through the warp's staging tile `stg` (32 rows x 80 B) so that every global access is a
// 64-byte row segment: lane l owns row l in the math, and serves rows (l >> 2) + 8 i, 16-byte column (l & 3) in the I/O.
constexpr int CONV_TAB_ENTRIES = 1024;                   // bias / gamma tables in shared memory (Cout <= 1024 per launch)
constexpr int CONV_TAB_BYTES = 2 * CONV_TAB_ENTRIES * 2;

// bias and gamma of the launch, zero-padded, copied once per CTA: the epilogue reads them per 32-column chunk, and a
// global (even L1-resident) load per chunk sat on its critical path.  Called by `nthreads` threads with ids 0..
__device__ __forceinline__ void conv_fill_tables(const ConvParams& p, __nv_bfloat16* tab, int tid, int nthreads) {
  for (int i = tid; i < CONV_TAB_ENTRIES; i += nthreads) {
    tab[i] = (p.bias && i < p.cout_store) ? p.bias[i] : __float2bfloat16_rn(0.0f);
    tab[CONV_TAB_ENTRIES + i] = (p.gamma && i < p.cout_store) ? p.gamma[i] : __float2bfloat16_rn(0.0f);
  }
}

// L2 prefetch of the residual rows of the coming tile, issued before the wait for its accumulator: the loads of the
// epilogue then hit L2 instead of exposing a DRAM round trip per 32-column chunk.
template <int NCH>
__device__ __forceinline__ void conv_prefetch_residual(const ConvParams& p, int n0, int lane,
                                                       const long long (&pos_co)[4], const bool (&ok_co)[4]) {
  if (p.R == nullptr || (lane & 3) != 0) return;
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    const int col0 = n0 + c * 32;
    if (c * 32 < p.BN && col0 < p.cout_store) {
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (ok_co[i]) asm volatile("prefetch.global.L2 [%0];" ::"l"(p.R + pos_co[i] * p.ldr + col0));
    }
  }
}

template <int NCH, bool kRemoteArrive = false>
__device__ __forceinline__ void conv_epilogue_rows(const ConvParams& p, uint32_t t_addr, int n0, uint8_t* stg,
                                                   const __nv_bfloat16* tab, int lane,
                                                   long long frame, long long pos_own, bool ok_own,
                                                   const long long (&pos_co)[4], const bool (&ok_co)[4],
                                                   uint32_t tempty_addr) {
  const int cv = lane & 3;
  uint32_t xs[NCH * 16];
  float ss = 0.0f;
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    const int col0 = n0 + c * 32;
    if (c * 32 < p.BN && col0 < p.cout_store) {
      uint32_t v[32];
      tmem_ld32(t_addr + c * 32, v);
      uint32_t rw[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) rw[j] = 0u;
      if (p.R) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          uint4 r4 = make_uint4(0, 0, 0, 0);
          if (ok_co[i] && col0 + cv * 8 < p.cout_store)
            r4 = *reinterpret_cast<const uint4*>(p.R + pos_co[i] * p.ldr + col0 + cv * 8);
          *reinterpret_cast<uint4*>(stg + ((lane >> 2) + 8 * i) * CONV_STG_PITCH + cv * 16) = r4;
        }
        __syncwarp();
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const uint4 r4 = *reinterpret_cast<const uint4*>(stg + lane * CONV_STG_PITCH + g * 16);
          rw[g * 4] = r4.x; rw[g * 4 + 1] = r4.y; rw[g * 4 + 2] = r4.z; rw[g * 4 + 3] = r4.w;
        }
        __syncwarp();
      }
      tmem_ld_wait();
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const int cg = col0 + g * 8;
        uint4 bv = make_uint4(0, 0, 0, 0);
        if (cg < p.cout_store) bv = *reinterpret_cast<const uint4*>(tab + cg);
        const uint32_t bw[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          // bf16(acc + bias), then the packed bf16 add of the residual (one rounding each, as the torch ops);
          // channels in [Cout, cout_store) are exact zeros (zero-filled weights, zero bias / residual padding)
          uint32_t xp = pack_bf16x2(__uint_as_float(v[g * 8 + 2 * j]) + bf16_lo(bw[j]),
                                    __uint_as_float(v[g * 8 + 2 * j + 1]) + bf16_hi(bw[j]));
          xp = hadd2_bf16(xp, rw[g * 4 + j]);        // rw == 0 without a residual: x + 0 is exact
          ss = fmaf(bf16_lo(xp), bf16_lo(xp), ss);
          ss = fmaf(bf16_hi(xp), bf16_hi(xp), ss);
          xs[c * 16 + g * 4 + j] = xp;
        }
      }
      if (p.Y) {
        if (p.ncthw) {
          if (c == 0 && ok_own) {       // (C, T, H, W) output: Cout <= 32 (checked on the host)
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const int ch = col0 + 2 * j;
              if (ch < p.Cout)
                p.Y[(long long)ch * p.To * frame + pos_own] =
                    __ushort_as_bfloat16((unsigned short)(xs[c * 16 + j] & 0xFFFFu));
              if (ch + 1 < p.Cout)
                p.Y[(long long)(ch + 1) * p.To * frame + pos_own] =
                    __ushort_as_bfloat16((unsigned short)(xs[c * 16 + j] >> 16));
            }
          }
        } else {
#pragma unroll
          for (int g = 0; g < 4; ++g)
            *reinterpret_cast<uint4*>(stg + lane * CONV_STG_PITCH + g * 16) =
                make_uint4(xs[c * 16 + g * 4], xs[c * 16 + g * 4 + 1], xs[c * 16 + g * 4 + 2],
                           xs[c * 16 + g * 4 + 3]);
          __syncwarp();
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            if (ok_co[i] && col0 + cv * 8 < p.cout_store)
              *reinterpret_cast<uint4*>(p.Y + pos_co[i] * p.ldy + col0 + cv * 8) =
                  *reinterpret_cast<const uint4*>(stg + ((lane >> 2) + 8 * i) * CONV_STG_PITCH + cv * 16);
          }
          __syncwarp();
        }
      }
    }
  }
  // accumulator fully read: hand it back before the second (register-only) pass
  tc_fence_before();
  __syncwarp();
  if (lane == 0) {
    if constexpr (kRemoteArrive) mbar_arrive_cluster(tempty_addr); else mbar_arrive(tempty_addr);
  }
  if (p.Y2) {
    const float rinv = sqrtf((float)p.Cout) / fmaxf(sqrtf(ss), 1e-12f);
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      const int col0 = n0 + c * 32;
      if (c * 32 < p.BN && col0 < p.cout_store) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const int cg = col0 + g * 8;
          uint4 gv = make_uint4(0, 0, 0, 0);
          if (cg < p.cout_store) gv = *reinterpret_cast<const uint4*>(tab + CONV_TAB_ENTRIES + cg);
          const uint32_t gw[4] = {gv.x, gv.y, gv.z, gv.w};
          uint32_t o[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float y0 = bf16_lo(xs[c * 16 + g * 4 + j]) * rinv * bf16_lo(gw[j]);
            float y1 = bf16_hi(xs[c * 16 + g * 4 + j]) * rinv * bf16_hi(gw[j]);
            // silu(y) = h tanh(h) + h, h = y / 2; without SiLU the tanh factor is replaced by 1 (h + h = y exactly)
            const float h0 = 0.5f * y0, h1 = 0.5f * y1;
            o[j] = pack_bf16x2(fmaf(h0, p.silu ? tanh_approx(h0) : 1.0f, h0), fmaf(h1, p.silu ? tanh_approx(h1) : 1.0f, h1));
          }
          *reinterpret_cast<uint4*>(stg + lane * CONV_STG_PITCH + g * 16) = make_uint4(o[0], o[1], o[2], o[3]);
        }
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          if (ok_co[i] && col0 + cv * 8 < p.cout_store)
            *reinterpret_cast<uint4*>(p.Y2 + pos_co[i] * p.ldy2 + col0 + cv * 8) =
                *reinterpret_cast<const uint4*>(stg + ((lane >> 2) + 8 * i) * CONV_STG_PITCH + cv * 16);
        }
        __syncwarp();
      }
    }
  }
}

template <int NCH>
__global__ void __launch_bounds__(CONV_THREADS, 1)
gf_conv3d_kernel(const __grid_constant__ ConvMaps maps, const ConvParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t stage_bytes = CONV_A_BYTES + (uint32_t)p.BN * 128u;
  const uint32_t stg_base = smem_base + p.stages * stage_bytes;
  const uint32_t bar_base = stg_base + 4 * CONV_STG_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (CONV_MAX_STAGES + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * CONV_MAX_STAGES + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * CONV_MAX_STAGES + 2 + a); };
  const uint32_t tmem_ptr_smem = bar_base + 8u * (2 * CONV_MAX_STAGES + 4);
  __nv_bfloat16* tab = reinterpret_cast<__nv_bfloat16*>(smem_raw + (bar_base + 256u - smem_u32(smem_raw)));
  auto smem_a = [&](int s) { return smem_base + s * stage_bytes; };
  auto smem_b = [&](int s) { return smem_base + s * stage_bytes + CONV_A_BYTES; };

  const int warp = threadIdx.x >> 5;
  const int tiles_per_frame = p.nWt * p.nHt;
  const int num_m_tiles = p.To * tiles_per_frame;
  const int num_tiles = num_m_tiles * p.num_n_tiles;
  const int num_cb = (p.Cin + CONV_BK - 1) / CONV_BK;
  const int BW = 1 << p.bw_shift, BH = CONV_BM >> p.bw_shift;

  if (warp == 0 && elect_one()) {
    for (int i = 0; i < 4; ++i) tma_prefetch_desc(&maps.a[i]);
    tma_prefetch_desc(&maps.b);
  }
  if (warp >= 2) conv_fill_tables(p, tab, (int)threadIdx.x - 64, CONV_THREADS - 64);
  if (warp == 1) {
    if (elect_one()) {
      for (int s = 0; s < p.stages; ++s) {
        mbar_init(full_bar(s), 1);
        mbar_init(empty_bar(s), 1);
      }
      for (int a = 0; a < 2; ++a) {
        mbar_init(tfull_bar(a), 1);
        mbar_init(tempty_bar(a), 4);
      }
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc<1>(tmem_ptr_smem, 512);
    tmem_relinquish<1>();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_ptr_smem));

  if (warp == 0) {
    // ===================================================== TMA producer
    if (elect_one()) {
      int stage = 0; uint32_t phase = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        const int mt = t / p.num_n_tiles, nt = t - mt * p.num_n_tiles;
        const int tt = mt / tiles_per_frame;
        const int rem = mt - tt * tiles_per_frame;
        const int ht = rem / p.nWt, wt = rem - ht * p.nWt;
        const int h0 = ht * BH, w0 = wt * BW, n0 = nt * p.BN;
        for (int tap = 0; tap < p.ntaps; ++tap) {
          const CUtensorMap* am = &maps.a[p.tap_map[tap]];
          const int ct = tt * p.st + p.tap_dt[tap], ch = h0 + p.tap_dh[tap], cw = w0 + p.tap_dw[tap];
          for (int cb = 0; cb < num_cb; ++cb) {
            mbar_wait(empty_bar(stage), phase ^ 1u);
            mbar_arrive_expect_tx(full_bar(stage), stage_bytes);
            tma_load_4d(smem_a(stage), am, full_bar(stage), cb * CONV_BK, cw, ch, ct);
            tma_load_2d(smem_b(stage), &maps.b, full_bar(stage), tap * p.Cin + cb * CONV_BK, n0);
            if (++stage == p.stages) { stage = 0; phase ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================== MMA issuer (whole warp walks the loop, one lane issues)
    {
      const uint32_t idesc = idesc_bf16(CONV_BM, (uint32_t)p.BN, 0, 0);
      constexpr uint64_t dbase = smem_desc_base(/*sbo=*/1024, /*lbo=*/16);
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * 256;
        uint32_t accum = 0;
        for (int tap = 0; tap < p.ntaps; ++tap) {
          for (int cb = 0; cb < num_cb; ++cb) {
            mbar_wait(full_bar(stage), phase);
            tc_fence_after();
            const uint64_t adesc0 = smem_desc(dbase, smem_a(stage)), bdesc0 = smem_desc(dbase, smem_b(stage));
            const int kvalid = min(CONV_BK, p.Cin - cb * CONV_BK);
            const int nk = (kvalid + 15) >> 4;
            if (elect_one()) {
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                if (k < nk)
                  umma_ss<1>(d_tmem, adesc0 + uint64_t(k * 2), bdesc0 + uint64_t(k * 2), idesc, (k == 0) ? accum : 1u);
              }
              tc_commit(empty_bar(stage));
            }
            __syncwarp();
            accum = 1u;
            if (++stage == p.stages) { stage = 0; phase ^= 1u; }
          }
        }
        if (elect_one()) tc_commit(tfull_bar(acc));
        __syncwarp();
        if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
      }
    }
  } else {
    // ===================================================== epilogue (warps 2..5)
    const int q = warp & 3;
    const int lane = (int)lane_id();
    const long long frame = (long long)p.Ho * p.Wo;
    uint8_t* stg = smem_raw + (stg_base - smem_u32(smem_raw)) + (warp - 2) * CONV_STG_BYTES;
    int acc = 0; uint32_t acc_phase = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      const int mt = t / p.num_n_tiles, nt = t - mt * p.num_n_tiles;
      const int tt = mt / tiles_per_frame;
      const int rem = mt - tt * tiles_per_frame;
      const int ht = rem / p.nWt, wt = rem - ht * p.nWt;
      const int r_own = q * 32 + lane;
      const int h_own = ht * BH + (r_own >> p.bw_shift), w_own = wt * BW + (r_own & (BW - 1));
      const bool ok_own = h_own < p.Ho && w_own < p.Wo;
      const long long pos_own = (long long)tt * frame + (long long)h_own * p.Wo + w_own;
      long long pos_co[4];
      bool ok_co[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int r = q * 32 + (lane >> 2) + 8 * i;
        const int h = ht * BH + (r >> p.bw_shift), w = wt * BW + (r & (BW - 1));
        ok_co[i] = h < p.Ho && w < p.Wo;
        pos_co[i] = (long long)tt * frame + (long long)h * p.Wo + w;
      }
      conv_prefetch_residual<NCH>(p, nt * p.BN, lane, pos_co, ok_co);
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      const uint32_t t_addr = tmem_base + (uint32_t(q * 32) << 16) + acc * 256;
      conv_epilogue_rows<NCH>(p, t_addr, nt * p.BN, stg, tab, lane, frame, pos_own, ok_own, pos_co, ok_co,
                              tempty_bar(acc));
      if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<1>(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// Halo form for 3x3 spatial windows (stride 1).
//
// The tap-by-tap form above moves 16 KB of A plus N x 128 B of weights into shared memory per k-block and reads them
// back for 4 MMAs: with 128 B/clk of shared-memory bandwidth shared by the operand reads and the TMA fill it is bound
// at 29 % (N = 96) to 60 % (N = 192) of the tensor rate.  Here a CTA owns a 16*MH x 8 patch of output positions and,
// per (dt, 64-channel block), loads ONE (16*MH + 2) x 10 halo of the input instead of nine shifted 128-row tiles:
// position (h, w) of the halo sits at row h*10 + w of a 128-byte-row, 128B-swizzled buffer, so the A operand of tap
// (dh, dw) for accumulator `mh` is the same buffer read through a descriptor that starts at row (mh*16 + dh)*10 + dw
// with a stride of 10 rows (1280 B) between 8-row groups -- the swizzle is a function of the absolute shared-memory
// address for TMA and the tensor core alike (checked on B200: descriptors with the 'matrix base offset' field set to
// (addr >> 7) & 7 give wrong results, plain start addresses are exact).
//
// kCG = 2: a CTA pair works on two horizontally adjacent patches with M = 256 MMAs; each CTA stages its own halo and
// half of the weight rows, all TMA bytes are accounted on the leader's barriers, the leader's MMA warp issues and its
// commits are multicast to both CTAs, the second CTA's epilogue warps arrive on the leader's barrier through the
// cluster address space (the protocol of gf_gemm.cu's pair mode).  A weight stage holds one window row (3 taps).
// Measured: N = 96 90.5 % tensor-active (1471 TFLOP/s alone), N = 192 99.9 % (1678 TFLOP/s).
//
// Threads: warp 0 TMA, warp 1 MMA (the whole warp walks the loop, one elected lane issues), 4*MH epilogue warps (one
// per TMEM lane quarter and accumulator).
constexpr int HALO_W = 8;                                // tile width in positions (= rows of a UMMA core matrix)
constexpr int HALO_MAX_A_STAGES = 4;
constexpr int HALO_MAX_B_STAGES = 6;

// MH = accumulators (16-row halves) per CTA: 2 for Cout <= 128 (two 128-column accumulators per TMEM stage), 1 for
// wider outputs (one accumulator of up to 256 columns per stage; Cout = 384 runs as two 192-column n-tiles).
template <int MH> struct HaloCfg {
  static constexpr int kTileH = 16 * MH;                                   // output rows per CTA tile
  static constexpr int kRows = (HALO_W + 2) * (kTileH + 2);                // 340 / 180 halo positions
  static constexpr int kTxBytes = kRows * 128;                             // 43,520 / 23,040
  static constexpr int kStageBytes = (kTxBytes + 1023) / 1024 * 1024;      // 44 KB / 23 KB
  static constexpr int kEpiWarps = 4 * MH;
  static constexpr int kThreads = 64 + 32 * kEpiWarps;                     // 320 / 192
};

template <int NCH, int kCG, int MH>
__global__ void __launch_bounds__(HaloCfg<MH>::kThreads, 1)
gf_conv3d_halo_kernel(const __grid_constant__ ConvMaps maps, const ConvParams p) {
  using Cfg = HaloCfg<MH>;
  constexpr int HALO_H = Cfg::kTileH;
  constexpr int HALO_TX_BYTES = Cfg::kTxBytes;
  constexpr int HALO_STAGE_BYTES = Cfg::kStageBytes;
  const int HALO_A_STAGES = p.a_stages;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t b_rows = (uint32_t)p.BN / kCG;            // weight rows staged by this CTA (pair: half of them each)
  const uint32_t b_tap_bytes = b_rows * 128u;              // weights of one tap: b_rows x 64 channels
  const uint32_t b_bytes = 3u * b_tap_bytes;               // a B stage holds one window row (dw = 0, 1, 2)
  const uint32_t b_base = smem_base + HALO_A_STAGES * HALO_STAGE_BYTES;
  const uint32_t stg_base = b_base + p.stages * b_bytes;
  const uint32_t bar_base = stg_base + Cfg::kEpiWarps * CONV_STG_BYTES;
  auto afull_bar = [&](int s) { return bar_base + 8u * s; };
  auto aempty_bar = [&](int s) { return bar_base + 8u * (4 + s); };
  auto bfull_bar = [&](int s) { return bar_base + 8u * (8 + s); };
  auto bempty_bar = [&](int s) { return bar_base + 8u * (8 + HALO_MAX_B_STAGES + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (8 + 2 * HALO_MAX_B_STAGES + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (10 + 2 * HALO_MAX_B_STAGES + a); };
  const uint32_t tmem_ptr_smem = bar_base + 8u * (12 + 2 * HALO_MAX_B_STAGES);
  __nv_bfloat16* tab = reinterpret_cast<__nv_bfloat16*>(smem_raw + (bar_base + 512u - smem_u32(smem_raw)));
  auto smem_a = [&](int s) { return smem_base + s * HALO_STAGE_BYTES; };
  auto smem_b = [&](int s) { return b_base + s * b_bytes; };

  const int warp = threadIdx.x >> 5;
  const uint32_t cta_rank = (kCG == 2) ? cluster_ctarank() : 0u;
  const bool leader = cta_rank == 0;
  const int cluster_id = blockIdx.x / kCG, num_clusters = gridDim.x / kCG;
  // work unit of a CTA (pair): kCG horizontally adjacent 32 x 8 tiles; a pair's odd tile past the frame edge is all
  // out-of-range positions (zero-filled loads, masked stores)
  const int units_per_row = (p.nWt + kCG - 1) / kCG;
  const int tiles_per_frame = units_per_row * p.nHt;
  const int num_tiles = p.To * tiles_per_frame * p.num_n_tiles;       // n-tile innermost: neighbours share the halo in L2
  const int num_cb = (p.Cin + CONV_BK - 1) / CONV_BK;
  const int kt = p.ntaps / 9;

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&maps.a[0]);
    tma_prefetch_desc(&maps.b);
  }
  if (warp >= 2) conv_fill_tables(p, tab, (int)threadIdx.x - 64, Cfg::kThreads - 64);
  if (warp == 1) {
    if (elect_one()) {
      for (int s = 0; s < HALO_A_STAGES; ++s) { mbar_init(afull_bar(s), 1); mbar_init(aempty_bar(s), 1); }
      for (int s = 0; s < p.stages; ++s) { mbar_init(bfull_bar(s), 1); mbar_init(bempty_bar(s), 1); }
      for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), Cfg::kEpiWarps * kCG); }
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc<kCG>(tmem_ptr_smem, 512);
    tmem_relinquish<kCG>();
  }
  tc_fence_before();
  if constexpr (kCG == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_ptr_smem));

  if (warp == 0) {
    // ===================================================== TMA producer
    if (elect_one()) {
      int as = 0; uint32_t aphase = 0;
      int bs = 0; uint32_t bphase = 0;
      for (int t = cluster_id; t < num_tiles; t += num_clusters) {
        const int mu = t / p.num_n_tiles, n0 = (t - mu * p.num_n_tiles) * p.BN;
        const int tt = mu / tiles_per_frame;
        const int rem = mu - tt * tiles_per_frame;
        const int ht = rem / units_per_row, wt = (rem - ht * units_per_row) * kCG + (int)cta_rank;
        const int h0 = ht * HALO_H + p.tap_dh[0], w0 = wt * HALO_W + p.tap_dw[0];   // tap 0 is (dh, dw) = (-ph, -pw)
        for (int dt = 0; dt < kt; ++dt) {
          const int ct = tt * p.st + p.tap_dt[dt * 9];
          for (int cb = 0; cb < num_cb; ++cb) {
            mbar_wait(aempty_bar(as), aphase ^ 1u);
            if constexpr (kCG == 1) {
              mbar_arrive_expect_tx(afull_bar(as), HALO_TX_BYTES);
              tma_load_4d(smem_a(as), &maps.a[0], afull_bar(as), cb * CONV_BK, w0, h0, ct);
            } else {
              // pair: each CTA streams its own halo, all bytes are accounted on the leader's barrier
              if (leader) mbar_arrive_expect_tx(afull_bar(as), 2 * HALO_TX_BYTES);
              tma_load_4d_cg2(smem_a(as), &maps.a[0], mapa(afull_bar(as), 0), cb * CONV_BK, w0, h0, ct);
            }
            if (++as == HALO_A_STAGES) { as = 0; aphase ^= 1u; }
            for (int dh = 0; dh < 3; ++dh) {
              mbar_wait(bempty_bar(bs), bphase ^ 1u);
              if constexpr (kCG == 1) {
                mbar_arrive_expect_tx(bfull_bar(bs), b_bytes);
#pragma unroll
                for (int dw = 0; dw < 3; ++dw)
                  tma_load_2d(smem_b(bs) + dw * b_tap_bytes, &maps.b, bfull_bar(bs),
                              (dt * 9 + dh * 3 + dw) * p.Cin + cb * CONV_BK, n0);
              } else {
                if (leader) mbar_arrive_expect_tx(bfull_bar(bs), 2 * b_bytes);
                const uint32_t lead_bar = mapa(bfull_bar(bs), 0);
#pragma unroll
                for (int dw = 0; dw < 3; ++dw)
                  tma_load_2d_cg2(smem_b(bs) + dw * b_tap_bytes, &maps.b, lead_bar,
                                  (dt * 9 + dh * 3 + dw) * p.Cin + cb * CONV_BK, n0 + (int)(cta_rank * b_rows));
              }
              if (++bs == p.stages) { bs = 0; bphase ^= 1u; }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================== MMA issuer
    // The whole warp walks the loop (barrier waits and descriptor arithmetic stay warp-uniform, i.e. in uniform
    // registers); one elected lane issues the MMAs and commits.  At N = 96 an MMA lasts 56 cycles, so the issue path
    // per MMA has to stay well below that: taps and k-steps are fully unrolled, descriptors differ in their low word
    // only (start address >> 4) and are built by adding constants.
    if (kCG == 1 || leader) {
      const uint32_t idesc = idesc_bf16(CONV_BM * kCG, (uint32_t)p.BN, 0, 0);
      constexpr uint64_t b_dbase = smem_desc_base(/*sbo=*/1024, /*lbo=*/16);
      constexpr uint64_t a_dbase = smem_desc_base(/*sbo=*/(HALO_W + 2) * 128, /*lbo=*/16);
      int as = 0; uint32_t aphase = 0;
      int bs = 0; uint32_t bphase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      for (int t = cluster_id; t < num_tiles; t += num_clusters) {
        mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * 256;
        uint32_t accum = 0;
        for (int dt = 0; dt < kt; ++dt) {
          for (int cb = 0; cb < num_cb; ++cb) {
            mbar_wait(afull_bar(as), aphase);
            tc_fence_after();
            const int kvalid = min(CONV_BK, p.Cin - cb * CONV_BK);
            const int nk = (kvalid + 15) >> 4;
            const uint32_t a_lo0 = (uint32_t)smem_desc(a_dbase, smem_a(as));
            constexpr uint32_t a_hi = (uint32_t)(a_dbase >> 32), b_hi = (uint32_t)(b_dbase >> 32);
            constexpr int kRowUnits = 128 / 16;                      // one halo row in descriptor address units
            constexpr int kHalf = 16 * (HALO_W + 2) * kRowUnits;     // second accumulator: 16 tile rows further down
#pragma unroll 1
            for (int dh = 0; dh < 3; ++dh) {
              mbar_wait(bfull_bar(bs), bphase);
              tc_fence_after();
              const uint32_t b_lo = (uint32_t)smem_desc(b_dbase, smem_b(bs));
              const uint32_t a_lo = a_lo0 + (uint32_t)(dh * (HALO_W + 2) * kRowUnits);
              const uint32_t b_tap_units = b_tap_bytes >> 4;
              if (elect_one()) {
                // one window row: 3 taps x MH accumulators x nk k-steps, straight-line for every nk
                auto issue = [&](auto nk_c) {
                  constexpr int NK = decltype(nk_c)::value;
#pragma unroll
                  for (int dw = 0; dw < 3; ++dw) {
#pragma unroll
                    for (int mh = 0; mh < MH; ++mh) {
#pragma unroll
                      for (int k = 0; k < NK; ++k)
                        umma_ss_lohi<kCG>(d_tmem + mh * 128, a_lo + dw * kRowUnits + mh * kHalf + k * 2, a_hi,
                                     b_lo + dw * b_tap_units + k * 2, b_hi, idesc, (dw == 0 && k == 0) ? accum : 1u);
                    }
                  }
                };
                switch (nk) {
                  case 4: issue(std::integral_constant<int, 4>{}); break;
                  case 3: issue(std::integral_constant<int, 3>{}); break;
                  case 2: issue(std::integral_constant<int, 2>{}); break;
                  default: issue(std::integral_constant<int, 1>{}); break;
                }
                tc_commit_group<kCG>(bempty_bar(bs));
              }
              __syncwarp();
              accum = 1u;
              if (++bs == p.stages) { bs = 0; bphase ^= 1u; }
            }
            if (elect_one()) tc_commit_group<kCG>(aempty_bar(as));
            __syncwarp();
            if (++as == HALO_A_STAGES) { as = 0; aphase ^= 1u; }
          }
        }
        if (elect_one()) tc_commit_group<kCG>(tfull_bar(acc));
        __syncwarp();
        if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
      }
    }
  } else {
    // ===================================================== epilogue (warps 2..9)
    const int ew = warp - 2;
    const int mh = ew >> 2;                     // accumulator (upper / lower 16 tile rows; always 0 for MH = 1)
    const int q = warp & 3;                     // TMEM lane quarter
    const int lane = (int)lane_id();
    const long long frame = (long long)p.Ho * p.Wo;
    uint8_t* stg = smem_raw + (stg_base - smem_u32(smem_raw)) + ew * CONV_STG_BYTES;
    int acc = 0; uint32_t acc_phase = 0;
    for (int t = cluster_id; t < num_tiles; t += num_clusters) {
      const int mu = t / p.num_n_tiles, n0 = (t - mu * p.num_n_tiles) * p.BN;
      const int tt = mu / tiles_per_frame;
      const int rem = mu - tt * tiles_per_frame;
      const int ht = rem / units_per_row, wt = (rem - ht * units_per_row) * kCG + (int)cta_rank;
      const int hbase = ht * HALO_H + mh * 16 + q * 4;       // tile row of this warp's row 0
      const int w0 = wt * HALO_W;
      // own row (TMEM lane) and the four rows this lane serves in the coalesced phase
      const int h_own = hbase + (lane >> 3), w_own = w0 + (lane & 7);
      const bool ok_own = h_own < p.Ho && w_own < p.Wo;
      const long long pos_own = (long long)tt * frame + (long long)h_own * p.Wo + w_own;
      long long pos_co[4];
      bool ok_co[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int row = (lane >> 2) + 8 * i;
        const int h = hbase + (row >> 3), w = w0 + (row & 7);
        ok_co[i] = h < p.Ho && w < p.Wo;
        pos_co[i] = (long long)tt * frame + (long long)h * p.Wo + w;
      }
      conv_prefetch_residual<NCH>(p, n0, lane, pos_co, ok_co);
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      const uint32_t t_addr = tmem_base + (uint32_t(q * 32) << 16) + acc * 256 + mh * 128;
      conv_epilogue_rows<NCH, kCG == 2>(p, t_addr, n0, stg, tab, lane, frame, pos_own, ok_own, pos_co, ok_co,
                                        kCG == 2 ? mapa(tempty_bar(acc), 0) : tempty_bar(acc));
      if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
    }
  }

  tc_fence_before();
  if constexpr (kCG == 2) cluster_sync_all(); else __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<kCG>(tmem_base, 512);
  }
}

template <int NCH>
static int launch_taps(const ConvMaps& maps, const ConvParams& p, int smem, int grid, cudaStream_t s) {
  auto kern = gf_conv3d_kernel<NCH>;
  static bool configured[64] = {};
  if (int e = gf_set_smem_once(configured, reinterpret_cast<const void*>(kern), 227 * 1024)) return e;
  kern<<<grid, CONV_THREADS, smem, s>>>(maps, p);
  return (int)cudaGetLastError();
}

template <int NCH, int kCG, int MH>
static int launch_halo(const ConvMaps& maps, const ConvParams& p, int smem, int clusters, cudaStream_t s) {
  auto kern = gf_conv3d_halo_kernel<NCH, kCG, MH>;
  static bool configured[64] = {};
  if (int e = gf_set_smem_once(configured, reinterpret_cast<const void*>(kern), 227 * 1024)) return e;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(clusters * kCG);
  cfg.blockDim = dim3(HaloCfg<MH>::kThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kCG;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return (int)cudaLaunchKernelEx(&cfg, kern, maps, p);
}

template <int kCG>
static int dispatch_halo(const ConvMaps& maps, const ConvParams& p, int smem, int clusters, cudaStream_t s) {
  switch (p.BN / 32) {
    case 1: return launch_halo<1, kCG, 2>(maps, p, smem, clusters, s);
    case 2: return launch_halo<2, kCG, 2>(maps, p, smem, clusters, s);
    case 3: return launch_halo<3, kCG, 2>(maps, p, smem, clusters, s);
    default: return launch_halo<4, kCG, 2>(maps, p, smem, clusters, s);
  }
}

// wide outputs (128 < BN <= 256): one accumulator per CTA, always as a CTA pair
static int dispatch_halo_wide(const ConvMaps& maps, const ConvParams& p, int smem, int clusters, cudaStream_t s) {
  return p.BN <= 192 ? launch_halo<6, 2, 1>(maps, p, smem, clusters, s) : launch_halo<8, 2, 1>(maps, p, smem, clusters, s);
}

static inline int floor_div(int a, int b) { return (a >= 0) ? a / b : -((-a + b - 1) / b); }

}  // namespace gf

extern "C" int gf_conv3d_cl_bf16(gf_ctx* ctx, const void* X, long long ldx, int T, int H, int W, int Cin, const void* Wt,
                                 int Cout, int kt, int kh, int kw, int st, int sh, int sw, int pt, int ph, int pw,
                                 const void* bias, void* Y, long long ldy, int To, int Ho, int Wo, const void* R,
                                 long long ldr, void* Y2, long long ldy2, const void* gamma, int silu, int out_ncthw,
                                 void* stream) {
  using namespace gf;
  if (!X || !Wt || (!Y && !Y2) || T <= 0 || H <= 0 || W <= 0 || To <= 0 || Ho <= 0 || Wo <= 0) return GF_ERR_BAD_ARG;
  if (Cin <= 0 || (Cin % 8) || (ldx % 8) || ldx <= 0 || Cout <= 0) return GF_ERR_BAD_ARG;   // ldx < Cin: overlapping windows
  if (kt < 1 || kh < 1 || kw < 1 || kt * kh * kw > CONV_MAX_TAPS) return GF_ERR_UNSUPPORTED;
  if (st < 1 || sh < 1 || sh > 2 || sw != sh) return GF_ERR_UNSUPPORTED;
  if (pt < 0 || ph < 0 || pw < 0 || pt > 8 || ph > 8 || pw > 8) return GF_ERR_BAD_ARG;
  if ((reinterpret_cast<uintptr_t>(X) | reinterpret_cast<uintptr_t>(Wt) | reinterpret_cast<uintptr_t>(Y) |
       reinterpret_cast<uintptr_t>(Y2) | reinterpret_cast<uintptr_t>(R) | reinterpret_cast<uintptr_t>(bias) |
       reinterpret_cast<uintptr_t>(gamma)) & 15)
    return GF_ERR_BAD_ARG;
  const int cout_store = (Cout + 7) & ~7;
  if (cout_store > CONV_TAB_ENTRIES) return GF_ERR_UNSUPPORTED;
  if (!out_ncthw && Y && ((ldy % 8) || ldy < cout_store)) return GF_ERR_BAD_ARG;
  if (out_ncthw && (R || Y2 || Cout > 32)) return GF_ERR_UNSUPPORTED;
  if (R && ((ldr % 8) || ldr < cout_store)) return GF_ERR_BAD_ARG;
  if (Y2 && (!gamma || (ldy2 % 8) || ldy2 < cout_store)) return GF_ERR_BAD_ARG;

  ConvParams p{};
  p.To = To; p.Ho = Ho; p.Wo = Wo;
  p.Cin = Cin; p.Cout = Cout; p.cout_store = cout_store;
  p.ntaps = kt * kh * kw;
  p.st = st;
  const int cout32 = (Cout + 31) & ~31;
  p.num_n_tiles = (cout32 + 255) / 256;
  p.BN = (((cout32 + p.num_n_tiles - 1) / p.num_n_tiles) + 31) & ~31;
  if (Y2 && p.num_n_tiles != 1) return GF_ERR_UNSUPPORTED;   // fused RMS_norm needs the whole channel row in one tile
  // M-tile patch: the BW x BH (= 128) shape that wastes the fewest rows on this frame size
  long long best = -1;
  for (int s = 3; s <= 7; ++s) {
    const int bw = 1 << s, bh = 128 >> s;
    const long long cover = (long long)((Wo + bw - 1) / bw) * bw * ((Ho + bh - 1) / bh) * bh;
    if (best < 0 || cover < best) { best = cover; p.bw_shift = s; }
  }
  const int BW = 1 << p.bw_shift, BH = 128 >> p.bw_shift;
  p.nWt = (Wo + BW - 1) / BW;
  p.nHt = (Ho + BH - 1) / BH;
  const int stage_bytes = CONV_A_BYTES + p.BN * 128;
  p.stages = CONV_SMEM_BUDGET / stage_bytes;
  if (p.stages > CONV_MAX_STAGES) p.stages = CONV_MAX_STAGES;
  p.ncthw = out_ncthw; p.silu = silu;
  p.Y = reinterpret_cast<__nv_bfloat16*>(Y); p.ldy = ldy;
  p.bias = reinterpret_cast<const __nv_bfloat16*>(bias);
  p.R = reinterpret_cast<const __nv_bfloat16*>(R); p.ldr = ldr;
  p.Y2 = reinterpret_cast<__nv_bfloat16*>(Y2); p.ldy2 = ldy2;
  p.gamma = reinterpret_cast<const __nv_bfloat16*>(gamma);

  // taps: input index = out*stride + d - pad = stride*(out + floor(e/stride)) + (e mod stride), e = d - pad
  int tap = 0;
  for (int dt = 0; dt < kt; ++dt)
    for (int dh = 0; dh < kh; ++dh)
      for (int dw = 0; dw < kw; ++dw, ++tap) {
        const int eh = dh - ph, ew = dw - pw;
        const int oh = floor_div(eh, sh), ow = floor_div(ew, sw);
        const int par_h = eh - oh * sh, par_w = ew - ow * sw;
        p.tap_map[tap] = (signed char)(par_h * sw + par_w);
        p.tap_dt[tap] = (signed char)(dt - pt);
        p.tap_dh[tap] = (signed char)oh;
        p.tap_dw[tap] = (signed char)ow;
      }

  const CtxTuning tune = gf_ctx_tuning(ctx);
  const bool halo_ok = kh == 3 && kw == 3 && sh == 1 && st == 1 && tune.conv_impl != 1;
  const bool wide = cout32 > 128;                       // one accumulator per CTA, pair only
  if (halo_ok && (!wide || (Wo > HALO_W && tune.conv_impl != 2))) {
    // narrow: 32 x 8 output positions per CTA (two accumulators); wide: 16 x 8 (one accumulator of <= 256 columns, Cout
    // split evenly into n-tiles).  One input halo per (dt, channel block); see gf_conv3d_halo_kernel.
    const int MH = wide ? 1 : 2;
    if (wide) {
      p.num_n_tiles = (cout32 + 255) / 256;
      p.BN = (((cout32 + p.num_n_tiles - 1) / p.num_n_tiles) + 31) & ~31;
      if (Y2 && p.num_n_tiles != 1) return GF_ERR_UNSUPPORTED;
    } else {
      p.num_n_tiles = 1;
      p.BN = cout32;
    }
    p.bw_shift = 3;
    p.nWt = (Wo + HALO_W - 1) / HALO_W;
    p.nHt = (Ho + 16 * MH - 1) / (16 * MH);
    // pair form (cta_group::2): two adjacent tiles per CTA pair, M = 256 MMAs, each CTA stages half of the weight
    // rows.  Shared-memory bandwidth (128 B/clk for operand reads AND the TMA fill) is what bounds these kernels: at
    // N = 96 a single CTA moves 7 KB + 2.8 KB per MMA (76 cycles measured against 48 of math), the pair 5.5 + 1.8 KB.
    const int cg = (!wide && (tune.conv_impl == 2 || p.nWt < 2)) ? 1 : 2;
    const int a_stage_bytes = wide ? HaloCfg<1>::kStageBytes : HaloCfg<2>::kStageBytes;
    const int b_stage_bytes = 3 * (p.BN / cg) * 128;                  // one window row (3 taps) of this CTA's weight rows
    p.a_stages = wide ? 3 : 2;
    const int fixed = p.a_stages * a_stage_bytes + 4 * MH * CONV_STG_BYTES + 1024 + 512 + CONV_TAB_BYTES;
    p.stages = (224 * 1024 - fixed) / b_stage_bytes;
    if (p.stages > HALO_MAX_B_STAGES) p.stages = HALO_MAX_B_STAGES;
    if (p.stages < 2) return GF_ERR_UNSUPPORTED;
    ConvMaps hm;
    const uint64_t dims[4] = {(uint64_t)Cin, (uint64_t)W, (uint64_t)H, (uint64_t)T};
    const uint64_t strides[3] = {(uint64_t)ldx * 2, (uint64_t)ldx * 2 * W, (uint64_t)ldx * 2 * W * H};
    const uint32_t box[4] = {(uint32_t)CONV_BK, (uint32_t)(HALO_W + 2), (uint32_t)(16 * MH + 2), 1u};
    int hrc = gf_make_tmap_4d_bf16(&hm.a[0], X, dims, strides, box);
    if (hrc) return hrc;
    for (int i = 1; i < 4; ++i) hm.a[i] = hm.a[0];
    hrc = gf_make_tmap_2d_bf16(&hm.b, Wt, (uint64_t)p.ntaps * Cin, (uint64_t)Cout, (uint64_t)p.ntaps * Cin, CONV_BK,
                               (uint32_t)(p.BN / cg));
    if (hrc) return hrc;
    const int hsmem = fixed + p.stages * b_stage_bytes;
    const long long units = (long long)To * p.nHt * ((p.nWt + cg - 1) / cg) * p.num_n_tiles;
    int clusters = gf_num_sms() / cg;
    if (clusters <= 0) return GF_ERR_NO_DRIVER;
    if (clusters > units) clusters = (int)units;
    cudaStream_t hs = reinterpret_cast<cudaStream_t>(stream);
    if (wide) return dispatch_halo_wide(hm, p, hsmem, clusters, hs);
    return cg == 1 ? dispatch_halo<1>(hm, p, hsmem, clusters, hs) : dispatch_halo<2>(hm, p, hsmem, clusters, hs);
  }

  ConvMaps maps;
  const char* xb = reinterpret_cast<const char*>(X);
  int rc = 0;
  for (int par_h = 0; par_h < sh; ++par_h)
    for (int par_w = 0; par_w < sw; ++par_w) {
      if (par_h >= H || par_w >= W) return GF_ERR_BAD_ARG;
      const uint64_t dims[4] = {(uint64_t)Cin, (uint64_t)((W - par_w + sw - 1) / sw), (uint64_t)((H - par_h + sh - 1) / sh),
                                (uint64_t)T};
      const uint64_t strides[3] = {(uint64_t)ldx * 2 * sw, (uint64_t)ldx * 2 * W * sh, (uint64_t)ldx * 2 * W * H};
      const uint32_t box[4] = {(uint32_t)CONV_BK, (uint32_t)BW, (uint32_t)BH, 1u};
      rc = gf_make_tmap_4d_bf16(&maps.a[par_h * sw + par_w], xb + ((long long)par_h * W + par_w) * ldx * 2, dims,
                                strides, box);
      if (rc) return rc;
    }
  for (int i = sh * sw; i < 4; ++i) maps.a[i] = maps.a[0];
  rc = gf_make_tmap_2d_bf16(&maps.b, Wt, (uint64_t)p.ntaps * Cin, (uint64_t)Cout, (uint64_t)p.ntaps * Cin, CONV_BK,
                            (uint32_t)p.BN);
  if (rc) return rc;

  const int smem = p.stages * stage_bytes + 4 * CONV_STG_BYTES + 1024 + 256 + CONV_TAB_BYTES;
  const long long tiles = (long long)To * p.nWt * p.nHt * p.num_n_tiles;
  int grid = gf_num_sms();
  if (grid <= 0) return GF_ERR_NO_DRIVER;
  if (grid > tiles) grid = (int)tiles;
  cudaStream_t cs = reinterpret_cast<cudaStream_t>(stream);
  return p.BN <= 128 ? launch_taps<4>(maps, p, smem, grid, cs) : launch_taps<8>(maps, p, smem, grid, cs);
}
