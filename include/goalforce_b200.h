/* goalforce_b200.h -- C ABI of libgoalforce_b200.so: the sm_100a kernels behind the Goal Force denoising hot path.
 *
 * The reference (brown-palm/goal-force) is pure Python/PyTorch and has no FFI of its own; every entry point below
 * replaces a span of torch ops in the reference and is what a maintainer would bind with ctypes.
 * Citations are relative to the reference checkout.
 *
 * Conventions
 *   - all tensors are device pointers to row-major bf16 unless stated; the caller owns every buffer (kernels never
 *     allocate); pointers must be 16-byte aligned and row pitches ("ld*", in elements) multiples of 8;
 *   - `stream` is a cudaStream_t passed as void*; calls are asynchronous and re-entrant per stream;
 *   - the library keeps no process-wide state: tuning and cached TMA descriptors live in an opaque `gf_ctx` that the
 *     tensor-core entry points take as their first argument (NULL = stateless call with default tuning); a context
 *     may be used by one thread at a time per call, its descriptor cache is internally locked;
 *   - the only allocations the library makes are the explicit gf_peer_alloc / gf_peer_status_alloc / gf_ctx_create
 *     objects; compute entry points never allocate;
 *   - return value: 0 on success, a GF_ERR_* code (< 0) for rejected arguments, or a positive cudaError_t.
 */
#ifndef GOALFORCE_B200_H_
#define GOALFORCE_B200_H_

#ifdef __cplusplus
extern "C" {
#endif

#define GF_ABI_VERSION 3

#define GF_ERR_BAD_ARG (-1)      /* null pointer, misaligned pointer/pitch, unsupported size */
#define GF_ERR_NO_DRIVER (-2)    /* cuTensorMapEncodeTiled not resolvable (no driver / no GPU) */
#define GF_ERR_TMAP (-3)         /* tensor-map encoding failed */
#define GF_ERR_UNSUPPORTED (-4)  /* shape outside what the kernels are built for */

/* GEMM epilogues (gf_gemm_bf16 `epi`) */
#define GF_MAX_PEERS 8           /* GPUs of one NVSwitch domain that can take part in a peer exchange */
#define GF_PEER_HANDLE_BYTES 64  /* size of an exported buffer handle (cudaIpcMemHandle_t) */

#define GF_EPI_BIAS 0       /* C = A.W^T + bias                                                */
#define GF_EPI_BIAS_GELU 1  /* C = gelu_tanh(A.W^T + bias)          wan_video_dit.py:209-210   */
#define GF_EPI_BIAS_SILU 2  /* C = silu(A.W^T + bias)               wan_video_dit.py:314-318   */
#define GF_EPI_GATE_RES 3   /* C = R + gate[n]*(A.W^T + bias)       wan_video_dit.py:189-194,226-229 */
#define GF_EPI_F32 4        /* C = A.W^T + bias stored as fp32 (ldc in floats)   wan_video_vae.py:325-337 */

int gf_abi_version(void);

/* Opaque per-caller context: TMA descriptor cache (keyed by address + geometry, so a descriptor is encoded once per
 * buffer instead of once per launch) and tuning.  gf_ctx_set_attention: impl 0 = per-shape choice (for long key
 * sequences the 80-row decoupled kernel, as a CTA pair when the 512-row work items fill the SM pairs; the 128-row
 * kernel for Lk <= 1024), 80 / 160 / 128 = forced (160 = CTA-pair form of 80); emu_pairs -1 = kernel default, or
 * 0/2/4/6 column pairs per 16 whose exp2 runs on the FMA pipe.  gf_ctx_set_gemm_raster: rasterisation group height in
 * m-tiles, 0 = per-shape choice.  gf_ctx_set_gemm_tile: tile width of the CTA-pair GEMM, 0 = per-shape choice (256, or
 * 224 where the 256-wide tiling would strand most of the last wave of SM pairs), 224 / 256 = forced.  gf_ctx_set_conv: impl 0 = per-shape choice of the convolution kernel (halo form for
 * 3x3 windows with Cout <= 128, as a CTA pair), 1 = always the tap-by-tap form, 2 = halo form on single CTAs.
 * gf_ctx_stats: descriptor-cache counters (any pointer may be NULL). */
typedef struct gf_ctx gf_ctx;
int gf_ctx_create(gf_ctx** ctx);
int gf_ctx_destroy(gf_ctx* ctx);
int gf_ctx_set_attention(gf_ctx* ctx, int impl, int emu_pairs);
int gf_ctx_set_gemm_raster(gf_ctx* ctx, int group_m);
int gf_ctx_set_gemm_tile(gf_ctx* ctx, int bn);
int gf_ctx_set_conv(gf_ctx* ctx, int impl);
int gf_ctx_stats(gf_ctx* ctx, long long* tmap_entries, long long* tmap_hits, long long* tmap_misses);

/* Number of SMs of the current CUDA device (148 on B200); <= 0 if no device. */
int gf_device_sms(void);

/* nn.Linear with fused tail: C[M,N] = epi(A[M,K] . W[N,K]^T).  tcgen05/TMEM/TMA GEMM.
 * Replaces F.linear at wan_video_dit.py:141-143,147,177-179,186,229, text/time MLPs :309-320, head :259 and the
 * k=1 Conv1d zero-conv at src/goal_force/wan_video_new.py:1564-1570.
 * bias/gate: [N] bf16 or NULL.  R: [M,ldr] residual for GF_EPI_GATE_RES (may alias C).  N % 32 == 0, K % 8 == 0.
 * cta_group: 1 = one CTA per 128x256 tile, 2 = CTA pair (cta_group::2) per 256x256 tile. */
int gf_gemm_bf16(gf_ctx* ctx, const void* A, long long lda, const void* W, long long ldw, void* C, long long ldc, int M, int N,
                 int K, const void* bias, int epi, const void* gate, const void* R, long long ldr, int cta_group,
                 void* stream);

/* LayerNorm over the last dim (fp32 statistics) followed by either adaLN modulation or an affine transform:
 *   weight == NULL : y = LN(x) * (1 + scale) + shift   (norm1/norm2 + modulate, wan_video_dit.py:64-65,225,228;
 *                                                        Head, :262-268).  shift/scale: [d] bf16.
 *   weight != NULL : y = LN(x) * weight + bias          (norm3, wan_video_dit.py:208,227)
 * x: [rows, ldx], y: [rows, ldy]; d % 256 == 0. */
int gf_layernorm_bf16(const void* x, long long ldx, void* y, long long ldy, int rows, int d, float eps,
                      const void* shift, const void* scale, const void* weight, const void* bias, void* stream);

/* In-place full-row RMSNorm (* weight) followed by interleaved-pair 3-D RoPE.
 * Replaces RMSNorm.forward + rope_apply (wan_video_dit.py:92-111,141-145,177-178).
 * x: [rows, ldx] (only the first d columns of each row are touched); weight: [d];
 * cos_sin: [rows, head_dim/2, 2] fp32 (cos, sin) table or NULL for no rotation (cross-attention q/k). */
int gf_rmsnorm_rope_bf16(void* x, long long ldx, int rows, int d, const void* weight, float eps,
                         const float* cos_sin, int head_dim, void* stream);

/* The q and k halves of a fused [rows, >= 2d] q|k|v row in one launch: columns [0,d) use weight_q, [d,2d) weight_k,
 * each with its own full-row statistic (SelfAttention.forward, wan_video_dit.py:141-145). Same math as
 * gf_rmsnorm_rope_bf16 applied twice. */
int gf_qk_rmsnorm_rope_bf16(void* qkv, long long ld, int rows, int d, const void* weight_q, const void* weight_k,
                            float eps, const float* cos_sin, int head_dim, void* stream);

/* Multi-head attention, no mask, no dropout: O = softmax(Q K^T * scale) V per head.  tcgen05 flash attention.
 * Replaces flash_attention() (wan_video_dit.py:28-61) for self-attention (Lk == Lq ~ 32k) and cross-attention
 * (Lk = 512).  Element (row, head, j) of Q lives at Q[row*ldq + head*head_dim + j]; same for K, V, O.
 * head_dim must be 128. */
int gf_attention_bf16(gf_ctx* ctx, const void* Q, long long ldq, const void* K, long long ldk, const void* V, long long ldv,
                      void* O, long long ldo, int Lq, int Lk, int heads, int head_dim, float scale, void* stream);

/* Patch gather for the (1,2,2) Conv3d patch embedding (wan_video_dit.py:307-308,341-349; ControlNet
 * src/goal_force/wan_video_new.py:83,91-92).  Reads channels of up to two NCTHW tensors (the torch.cat([x, y]) at
 * src/goal_force/wan_video_new.py:1457-1458 is never materialised) and writes tokens
 *   out[(f*H2 + h)*W2 + w, c*4 + kh*2 + kw] = src[c, f, 2h+kh, 2w+kw],   H2 = H/2, W2 = W/2
 * which is the A operand of the patch-embedding GEMM against weight.view(dim, C*4). src1 may be NULL (C1 = 0). */
int gf_patch_gather_bf16(const void* src0, int C0, const void* src1, int C1, void* out, long long ldo, int F, int H,
                         int W, void* stream);

/* unpatchify (wan_video_dit.py:351-356): head output [L, 4*C] with column (kh*2+kw)*C + c  ->  (C, F, H, W). */
int gf_unpatchify_bf16(const void* tokens, long long ldt, void* out, int C, int F, int H, int W, void* stream);

/* y[i, :] = a[i, :] + b[:]  (bf16 add with one rounding; modulation + t_mod, wan_video_dit.py:218-219,264-267). */
int gf_add_rows_bf16(const void* a, const void* b, void* y, int rows, int cols, void* stream);

/* y = a + b elementwise, n bf16 elements, one rounding; y may alias a or b.  The strided ControlNet inject
 * `x = x + states[i // stride]` (src/goal_force/wan_video_new.py:1559-1563). */
int gf_add_bf16(const void* a, const void* b, void* y, long long n, void* stream);

/* y = silu(x) elementwise (time_projection.0, wan_video_dit.py:319-320). */
int gf_silu_bf16(const void* x, void* y, long long n, void* stream);

/* Classifier-free guidance + flow-matching Euler step in one pass
 * (src/goal_force/wan_video_new.py:716,721; diffsynth/schedulers/flow_match.py:72-82):
 *   pred = nega + cfg_scale * (posi - nega)      (nega == NULL: pred = posi)
 *   latents_out = latents + pred * dsigma         dsigma = sigma_next - sigma
 * with bf16 rounding after every torch op of the reference. latents_out may alias latents. */
int gf_cfg_euler_bf16(const void* posi, const void* nega, const void* latents, void* latents_out, float cfg_scale,
                      float dsigma, long long n, void* stream);

/* sinusoidal_embedding_1d (wan_video_dit.py:68-72): out[b, :] = [cos(t_b w_i) | sin(t_b w_i)], w_i = 10000^(-i/(dim/2)),
 * angles in float64, result rounded to bf16. timestep: [B] bf16 on the device (no host sync). */
int gf_timestep_embedding_bf16(const void* timestep, void* out, int B, int dim, void* stream);

/* ---- umT5 prompt encoder pieces (diffsynth/models/wan_video_text_encoder.py; its linears are gf_gemm_bf16) ----------
 * gf_embedding_bf16: out[r, :] = table[ids[r], :] (token_embedding, :236); ids are int64 on the device.
 * gf_t5_rmsnorm_bf16: T5LayerNorm (:18-30), y = bf16(x * rsqrt(mean(x^2) + eps)) * weight, out of place, d % 8 == 0.
 * gf_mul_bf16: y = a * b elementwise (fc1(x) * gelu(gate(x)), :95-100); y may alias a or b.
 * gf_t5_attention_bf16: T5Attention.forward (:47-78) for `batch` sequences: softmax(q k^T + pos_bias + mask) v per head,
 *   no 1/sqrt(d) scaling, head_dim 64, Lk <= 512.  Row b*L + i of Q/K/V/O is token i of sequence b, head h at columns
 *   [64 h, 64 h + 64).  pos_bias[h, i, j] = bias_table[bucket_of[j - i + Lq - 1] * heads + h] with bias_table the
 *   [num_buckets, heads] bf16 weight of T5RelativeEmbedding (:136-175) and bucket_of the int32 [Lq + Lk - 1] bucket
 *   index per relative position (host-built with the reference's formula); key_mask: [batch, Lk] int32, 0 = padding
 *   (filled with finfo(bf16).min like :66-70), or NULL. */
int gf_embedding_bf16(const long long* ids, const void* table, void* out, int rows, int dim, long long vocab,
                      void* stream);
int gf_t5_rmsnorm_bf16(const void* x, long long ldx, void* y, long long ldy, int rows, int d, const void* weight,
                       float eps, void* stream);
int gf_mul_bf16(const void* a, const void* b, void* y, long long n, void* stream);
int gf_t5_attention_bf16(const void* Q, long long ldq, const void* K, long long ldk, const void* V, long long ldv,
                         void* O, long long ldo, int batch, int Lq, int Lk, int heads, int head_dim,
                         const void* bias_table, const int* bucket_of, const int* key_mask, void* stream);

/* ---- Wan video VAE (diffsynth/models/wan_video_vae.py) -------------------------------------------------
 * Clips are channels-last: X[t][h][w][c] bf16 with position pitch ld* (elements, multiple of 8).
 *
 * gf_conv3d_cl_bf16: 3-D convolution as an implicit GEMM on the tensor cores (CausalConv3d :33-52, Resample's Conv2d
 * :92-119 with kt = 1, the strided time_conv :104-119).
 *   Y[to,ho,wo,co] = sum Wt[co][(dt,dh,dw)][ci] * X[to*st + dt - pt, ho*sh + dh - ph, wo*sw + dw - pw, ci]  (+ bias)
 * Input positions outside the clip read as zero, so pt/ph/pw are the leading paddings (pt = kt-1 is the causal form)
 * and the trailing padding follows from To/Ho/Wo.  Wt: [Cout][kt*kh*kw][Cin] bf16, Cin % 8 == 0; sh == sw in {1, 2}.
 * ldx < Cin is allowed: every position then sees a window of Cin elements that overlaps its right-hand neighbours
 * (the 3-channel input convolution folds its three dw taps into one 64-element window this way).
 * R (optional): residual [To*Ho*Wo][ldr] added after the bias (ResidualBlock :296-301).
 * Y2/gamma (optional, Cout <= 256): Y2 = act(RMS_norm(Y) * gamma) of the NEXT layer (:55-70), act = SiLU when silu != 0;
 * Y may be NULL when only Y2 is wanted.  out_ncthw != 0: Y is (Cout, To, Ho, Wo) planes (the 3-channel decoder head).
 * bias/gamma hold at least Cout rounded up to 8 elements. */
int gf_conv3d_cl_bf16(gf_ctx* ctx, const void* X, long long ldx, int T, int H, int W, int Cin, const void* Wt, int Cout,
                      int kt, int kh, int kw, int st, int sh, int sw, int pt, int ph, int pw, const void* bias, void* Y,
                      long long ldy, int To, int Ho, int Wo, const void* R, long long ldr, void* Y2, long long ldy2,
                      const void* gamma, int silu, int out_ncthw, void* stream);

/* RMS_norm over the channels of every position (:55-70): y = x / max(|x|_2, 1e-12) * sqrt(C) * gamma, then SiLU when
 * silu != 0 (ResidualBlock :283-288, heads :581,801).  C % 8 == 0, C <= 1024. */
int gf_vae_rmsnorm_bf16(const void* x, long long ldx, void* y, long long ldy, long long rows, int C, const void* gamma,
                        int silu, void* stream);

/* Nearest-exact 2x spatial upsampling of F frames (Upsample :74-80): out[f][2h+a][2w+b] = src_f[h][w].
 * rest == NULL: src_f = first[f].  rest != NULL (upsample3d :138-160): output frame 0 reads first[0]; output frame
 * f >= 1 reads rest[(f-1)/2] at channel offset ((f-1)%2)*C, where rest is the [.,.,.,2C] time_conv result. */
int gf_vae_upsample2x_bf16(const void* first, long long ld_first, const void* rest, long long ld_rest, void* out,
                           long long ldo, int F, int H, int W, int C, void* stream);

/* P[r, :L] = softmax(S[r, :L] * scale) in bf16, P[r, L:Lp] = 0; S fp32 (gf_gemm_bf16 with GF_EPI_F32).  The
 * single-head attention of AttentionBlock (:304-342) is S = q k^T, this, then P v as GEMMs. */
int gf_softmax_f32_bf16(const float* S, long long lds, void* P, long long ldp, int rows, int L, int Lp, float scale,
                        void* stream);

/* (C, N) planes <-> channels-last rows.  planes_to_cl pads channels [C, Cp) with zeros; mode 1 applies the latent
 * un-normalisation z / inv_std + mean of VideoVAE_.decode (:1014-1018).  cl_to_planes mode 1 applies
 * (mu - mean) * inv_std of VideoVAE_.encode (:1003-1007).  mean / inv_std: fp32 [C] on the device.
 * row_w > 0: the destination has rows of row_w + 2*wpad positions and source position n lands at column n % row_w + wpad
 * of row n / row_w (the border positions are left untouched: the caller zero-fills them once). */
int gf_vae_planes_to_cl_bf16(const void* src, long long N, int C, void* dst, long long ldo, int Cp, const float* mean,
                             const float* inv_std, int mode, int row_w, int wpad, void* stream);
int gf_vae_cl_to_planes_bf16(const void* src, long long ld, long long N, int C, void* dst, const float* mean,
                             const float* inv_std, int mode, void* stream);

/* Decoder head (3x3x3 convolution to C <= 4 channels, :801-803) in two steps: gf_conv3d_cl_bf16 with a (3,1,1) kernel
 * whose output channel tap*4 + co holds the partial sum of spatial tap (dh, dw) = (tap / 3, tap % 3), then this gather:
 *   out[co, t, h, w] = bias[co] + sum_tap D[t, h + dh - 1, w + dw - 1, tap*4 + co]       (zero outside the frame)
 * D: [T*H*W][ld] bf16 (ld >= 36), bias: fp32 [C] on the device, out: (C, T, H, W) bf16. */
int gf_vae_head_gather_bf16(const void* D, long long ld, const float* bias, void* out, int C, int T, int H, int W,
                            void* stream);

/* Tile blending of WanVideoVAE.tiled_decode / tiled_encode (:1133-1153,1184-1204).
 * blend: values[c][t][h0+y][w0+x] += tile[c][t][y][x] * mask[y][x] (bf16 rounding after the product and after the sum).
 * blend_finish: values /= weight[h][w] (NULL: skip) and clamp to [-1, 1] when clamp != 0 (single_decode :1214-1217). */
int gf_vae_blend_bf16(void* values, int C, int T, int H, int W, const void* tile, int th, int tw, int h0, int w0,
                      const void* mask, void* stream);
int gf_vae_blend_finish_bf16(void* values, long long planes, int H, int W, const void* weight, int clamp, void* stream);

/* Ulysses layout helpers (replace the head<->sequence reshuffles inside xfuser's long-context attention called at
 * diffsynth/distributed/xdit_context_parallel.py:121-126).
 * pack:   x[rows, heads*head_dim] (pitch ldx) -> out[P][rows][ldo], destination rank p receives heads
 *         [p*heads/P, (p+1)*heads/P) in the first heads/P*head_dim elements of each out row (ldo lets q, k, v share
 *         one send buffer).
 * unpack: in[P][rows][heads/P*head_dim] (source-rank major) -> y[rows, heads*head_dim] (pitch ldy). */
int gf_ulysses_pack_bf16(const void* x, long long ldx, void* out, long long ldo, int rows, int heads, int head_dim,
                         int P, void* stream);
int gf_ulysses_unpack_bf16(const void* in, void* y, long long ldy, int rows, int heads, int head_dim, int P,
                           void* stream);

/* ---- fused Ulysses exchange over peer memory (NVLink / NVSwitch), one process per GPU --------------------------------
 * Replaces the all-to-all pair around self-attention that the reference delegates to xfuser
 * (diffsynth/distributed/xdit_context_parallel.py:110-131): the producing kernels store directly into the consumer
 * GPU's buffers, and a flag barrier on the compute stream replaces the collective.
 *
 * gf_peer_alloc: zero-filled device buffer that other processes of the node may map (cudaMalloc + CUDA IPC).
 * gf_peer_export / gf_peer_import: 64-byte handle of such a buffer / mapping of a peer's handle into this process
 * (the host side exchanges the handles, e.g. with torch.distributed.all_gather_object).
 * gf_peer_barrier: flag_peers[r] = rank r's flag buffer (>= 256 bytes, zero at start; own buffer at [rank]).
 * Enqueues a kernel that publishes this rank's earlier writes, signals the next epoch to every peer and waits until
 * every peer has signalled it.  The epoch counter lives in the flag buffer (so the call can be captured in a CUDA
 * graph); all ranks must issue the same sequence of barriers.  The wait is bounded by wall-clock time: if a peer
 * does not arrive within timeout_ms (<= 0: 60 s) the kernel stores 1 + peer index into *status (device pointer from
 * gf_peer_status_alloc, may be NULL) and returns -- no trap; the host polls the mapped word and raises.
 * gf_peer_status_alloc: 64 bytes of zeroed host-mapped memory; *host_ptr for the CPU, *dev_ptr for the kernels. */
int gf_peer_alloc(void** ptr, long long bytes);
int gf_peer_free(void* ptr);
int gf_peer_export(void* ptr, void* handle64);
int gf_peer_import(const void* handle64, void** ptr);
int gf_peer_unimport(void* ptr);
int gf_peer_barrier(void* const* flag_peers, int n_peers, int rank, long long timeout_ms, unsigned* status,
                    void* stream);
int gf_peer_status_alloc(unsigned** host_ptr, unsigned** dev_ptr);
int gf_peer_status_free(unsigned* host_ptr);

/* q/k RMSNorm + RoPE (as gf_qk_rmsnorm_rope_bf16) and v pass-through, written into the Ulysses receive buffers:
 * rank p owns heads [p*heads/P, (p+1)*heads/P), w = heads/P*head_dim columns.  Row r of this rank (global token
 * rank*rows + r) lands in recv_peers[p] at row rank*rows + r as [q_p | k_p | v_p] (3*w columns, pitch ld_recv).
 * qkv itself is not modified.  recv_peers: host array of n_peers device pointers (own buffer at [rank]). */
int gf_qkv_rmsnorm_rope_scatter_bf16(const void* qkv, long long ld, int rows, int d, const void* weight_q,
                                     const void* weight_k, float eps, const float* cos_sin, int head_dim,
                                     void* const* recv_peers, int n_peers, int rank, long long ld_recv, void* stream);

/* gf_attention_bf16 whose output rows go back to their owners: global query row g belongs to rank g / rows_per_peer
 * and is stored at O_peers[g / rows_per_peer] + (g % rows_per_peer)*ldo + col_offset + head*head_dim
 * (col_offset = rank * heads * head_dim: this rank's head group inside the owner's [rows, all heads] buffer). */
int gf_attention_scatter_bf16(gf_ctx* ctx, const void* Q, long long ldq, const void* K, long long ldk, const void* V, long long ldv,
                              void* const* O_peers, int n_peers, long long ldo, int rows_per_peer, int col_offset,
                              int Lq, int Lk, int heads, int head_dim, float scale, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GOALFORCE_B200_H_ */
