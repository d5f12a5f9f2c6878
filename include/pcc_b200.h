/*
 * pcc_b200.h -- C ABI of libpcc_b200.so, the B200 (sm_100a) implementation of the geometric hot path of
 * rhmes/point-cloud-compression.
 *
 * The reference has no FFI for this path: it calls plain Python functions (its own pn_kit helpers and the
 * PyTorch3D ops it imports).  Each entry point below replaces one of those calls; the Python host code in
 * point-cloud-compression_b200/pcc_b200/ binds them with ctypes and keeps the reference's signatures
 * (INTEGRATION.md shows the binding).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer to a dense, contiguous, caller-owned buffer (float32 / int64);
 *     the library allocates nothing persistent and never synchronises the host (except pcc_fps_f32 for
 *     clouds that need a cooperative launch, which only queries occupancy once per process);
 *   - `stream` is a cudaStream_t passed as void* (NULL = default stream); kernels are only enqueued;
 *   - return 0 on success, a negative PCC_ERR_* for an argument error, a positive cudaError_t if a launch
 *     failed; pcc_last_error_string() describes the last failure on the calling thread;
 *   - no C++ exception crosses this boundary, and there is NO CPU fallback: unsupported shapes are errors.
 *   - squared distances follow the reference CPU arithmetic bit for bit:
 *         d2 = fl(fl(fl(dx*dx) + fl(dy*dy)) + fl(dz*dz))      (no FMA contraction)
 *     and every tie goes to the lowest index.
 */
#ifndef PCC_B200_H
#define PCC_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PCC_ERR_INVALID_ARGUMENT (-1)
#define PCC_ERR_UNSUPPORTED (-2)

#define PCC_MAX_KNN_K 1024

/* Library version (major*10000 + minor*100 + patch) and the last error text of the calling thread. */
int pcc_version(void);
const char *pcc_last_error_string(void);
/* Number of kernels this library has enqueued so far in the process (every launch is counted once). */
int64_t pcc_launch_count(void);

/*
 * Farthest point sampling of `npoint` indices from each of B clouds of N points.
 * Replaces pn_kit.farthest_point_sample_batch (/root/reference/pn_kit.py:309-330): pass the reference's
 * torch.randint draw in start_idx[B] (device int64; values outside [0, N) are clamped into the cloud) and init_dist = 1e10;
 * and pytorch3d.ops.sample_farthest_points (/root/reference/pointnet_sa_module.py:10-13): start_idx = NULL
 * (start at index 0), init_dist = FLT_MAX.
 * out_idx[B, npoint] int64; entries k >= N are -1 (PyTorch3D padding when npoint > N).
 * out_xyz (nullable) [B, npoint, 3]: the sampled points themselves -- the index_points / gather call that follows every
 * reference FPS call (compress.py:96, pointnet_sa_module.py:68) -- snapped to the octree grid when quant_cube > 0
 * (floor(c / cube) * cube + cube / 2, octree_np.py:114-133); padded entries are 0.
 * Clouds with N > 8192 need a device `workspace` of pcc_fps_workspace_bytes(B, N, npoint) bytes (0 for N <= 8192 -> NULL
 * allowed): they are spread over several co-resident CTAs; long samplings (npoint >= ~N / 1024) of clouds of 196,608 to
 * 1,048,575 points (the 1M-point scene) continue, after a head of such iterations, on one CTA per cloud that keeps the cloud
 * sorted along a Hilbert curve in buckets and skips, exactly, every bucket the new centre cannot reach (about 25 bytes of
 * workspace per point) -- the same indices either way.
 */
int64_t pcc_fps_workspace_bytes(int B, int N, int npoint);
int pcc_fps_f32(const float *xyz, int B, int N, int npoint, const int64_t *start_idx, float init_dist,
                int64_t *out_idx, float *out_xyz, float quant_cube, void *workspace, void *stream);

/*
 * K nearest neighbours of every q[b, i] among p[b, :], ascending in (d2, idx).
 * Replaces pytorch3d.ops.knn_points (call sites /root/reference/train.py:185, compress.py:71, pn_kit.py:190,
 * pppe_pcd_ae.py:599, eval.py:132).  1 <= K <= PCC_MAX_KNN_K.
 *   out_d2 [B,P1,K] float32, out_idx [B,P1,K] int64 (either may be NULL when only out_nn is wanted);
 *   slots k >= P2 hold d2 = 0, idx = 0.
 *   out_nn (nullable) [B,P1,K,3]: the gathered neighbours (return_nn=True); with centre_sub != 0 the query is
 *   subtracted (the `grouped_xyz -= centre` that follows every reference call: train.py:188, compress.py:72,
 *   pn_kit.py:191) and the result is multiplied by nn_scale (train.py:192, compress.py:108; pass 1.0f to
 *   leave it out -- the multiply is a separately rounded fp32 op, as in the reference).
 */
int pcc_knn_f32(const float *q, const float *p, int B, int P1, int P2, int K, float *out_d2, int64_t *out_idx,
                float *out_nn, int centre_sub, float nn_scale, void *stream);

/*
 * Ball query: the first K indices j (in index order) with d2(q[b,i], p[b,j]) < radius*radius.
 * Replaces pytorch3d.ops.ball_query (/root/reference/pointnet_sa_module.py:16-19).
 *   out_idx [B,P1,K] int64 padded with -1; out_d2 (nullable) [B,P1,K] padded with 0.
 */
int pcc_ball_query_f32(const float *q, const float *p, int B, int P1, int P2, int K, float radius,
                       int64_t *out_idx, float *out_d2, void *stream);

/*
 * Row gather out[b, m, :] = feat[b, idx[b, m], :] for feat [B,N,C], idx [B,M] (any [B,S] / [B,S,K] index
 * tensor flattened).  Replaces pn_kit.index_points (/root/reference/pn_kit.py:332-360), pytorch3d knn_gather
 * (/root/reference/pointnet_sa_module.py:28) and the torch.gather at pointnet_sa_module.py:68.
 * Indices must lie in [0, N); negative indices are an error the caller clamps away first, as the reference does
 * (pointnet_sa_module.py:27).  pcc_gather_bwd_f32 scatter-adds grad_out into grad_feat (zero-filled by caller).
 */
int pcc_gather_f32(const float *feat, const int64_t *idx, int B, int N, int C, int64_t M, float *out,
                   void *stream);
int pcc_gather_bwd_f32(const float *grad_out, const int64_t *idx, int B, int N, int C, int64_t M,
                       float *grad_feat, void *stream);

/*
 * Nearest neighbour (K = 1) of every q[b, i] in p[b, :]: out_d2 [B,P1], out_idx (nullable) [B,P1] int64.
 * The inner step of the D1 PSNR loop (/root/reference/eval.py:68-81) and of chamfer_distance.
 * `workspace` must hold pcc_nn1_workspace_bytes(B, P1, P2) bytes (may be 0 -> NULL allowed).
 */
int64_t pcc_nn1_workspace_bytes(int B, int P1, int P2);
int pcc_nn1_f32(const float *q, const float *p, int B, int P1, int P2, float *out_d2, int64_t *out_idx,
                void *workspace, void *stream);

/*
 * Chamfer distance, both directions, point_reduction = batch_reduction = "mean".
 * Replaces pytorch3d.loss.chamfer_distance (/root/reference/AE.py:67, PPPF_AE.py:168, pppe_pcd_ae.py:820,
 * eval.py:204).
 *   out_dx [B,P1], out_ix [B,P1] (nullable), out_dy [B,P2], out_iy [B,P2] (nullable): per-point minima
 *   (bit-exact) and their argmins (needed by the backward);
 *   out_per_cloud [B] float32: mean_i dx + mean_j dy of each cloud;  out_loss [1] float32: their mean.
 * `workspace` must hold pcc_chamfer_workspace_bytes(B, P1, P2) bytes.
 */
int64_t pcc_chamfer_workspace_bytes(int B, int P1, int P2);
int pcc_chamfer_fwd_f32(const float *x, const float *y, int B, int P1, int P2, float *out_dx, int64_t *out_ix,
                        float *out_dy, int64_t *out_iy, float *out_per_cloud, float *out_loss, void *workspace,
                        void *stream);

/*
 * Chamfer backward (PyTorch3D knn_points_backward through both K=1 searches):
 *   gx[b,i] += 2*wx*(x_i - y_ix[i]),  gy[b,ix[i]] -= same   with wx = grad_loss / (B*P1), and symmetrically
 *   with wy = grad_loss / (B*P2).   grad_loss is a DEVICE scalar.  gx, gy are overwritten (zeroed inside).
 */
int pcc_chamfer_bwd_f32(const float *x, const float *y, const int64_t *ix, const int64_t *iy, int B, int P1,
                        int P2, const float *grad_loss, float *gx, float *gy, void *stream);

/*
 * Element-wise glue of the batched codec driver.
 * pcc_normalize_f32: pn_kit.normalize (/root/reference/pn_kit.py:47-60) per cloud: out = (p - center) * (1 - margin) /
 *   longest + 0.5; center [B,3], longest [B], bbox [B,6] (min xyz, max xyz) are optional outputs.
 * pcc_assemble_f32: decompress.py:104-116 -- patches [B*S,k,3] / patch_scale + centres [B,S,3], then pn_kit.denormalize
 *   with center / longest (both NULL: skip the de-normalisation); out [B, S*k, 3].
 */
int pcc_normalize_f32(const float *xyz, int B, int N, float margin, float *out, float *center, float *longest, float *bbox,
                      void *stream);
int pcc_assemble_f32(const float *patches, const float *centres, const float *center, const float *longest, int B, int S,
                     int k, float patch_scale, float margin, float *out, void *stream);

/*
 * Fused shared-MLP chain on the tensor cores (tcgen05, bf16 operands, fp32 accumulation):
 *     y = act_L(W_L . ... act_1(W_1 . x + b_1) ... + b_L)      per row of x [rows, ldx] (first cin columns used),
 * optionally followed by a max over every run of `group` consecutive rows (group <= 1: none).
 * Replaces the 1x1-conv stacks (+ max-pool) of pn_kit.SetAbstraction / PointNet / MLP
 * (/root/reference/pn_kit.py:196-207, 124-144, 289-305) and PointnetSAModule.mlp
 * (/root/reference/pointnet_sa_module.py:87-91).  Intermediate activations stay in shared memory / TMEM.
 *   - weights are packed once per layer with pcc_mlp_pack_weights_f32 from the reference's [cout, cin] fp32 tensor
 *     (Conv2d weight flattened) and [cout] bias into pcc_mlp_packed_bytes(cin, cout) bytes of device memory (bf16; the
 *     bias becomes an extra K column that multiplies a constant-one input channel, so it is rounded to bf16 too);
 *   - `layers` is a HOST array; every layer's weights must fit in shared memory together (otherwise
 *     PCC_ERR_UNSUPPORTED, never a fallback); group must divide 32, or be a multiple of 32 dividing 128, or be a
 *     multiple of 128 (then cout_L <= 128), and rows % group == 0;
 *   - out: fp32 [rows, cout_L] or [rows / group, cout_L], channel-last.
 * Numerics: operands rounded to bf16, products accumulated in fp32 (tolerance stated in tests/test_gpu_mlp.py).
 */
#define PCC_MLP_MAX_LAYERS 6
typedef struct PccMlpLayer {
    const void *packed_w; /* device, from pcc_mlp_pack_weights_f32 (weights and bias) */
    int cin, cout;
    int relu;             /* apply max(x, 0) after this layer */
    const float *w_f32;   /* optional: the raw [cout, cin] fp32 weights and [cout] bias (device).  When layer 0 has  */
    const float *b_f32;   /* cin <= 8 and these are given, it runs in fp32 on the CUDA cores (inputs not rounded).   */
} PccMlpLayer;
/* One input segment: rows of `channels` values taken from ptr[(row / row_div) * ld + 0..channels), fp32 (dtype 0) or
 * bf16 (dtype 1).  Segments are concatenated along the channel axis in order (the torch.cat of AE.py:39,51 without
 * materialising it); row_div > 1 broadcasts one source row over row_div consecutive positions (AE.py:50). */
#define PCC_MLP_MAX_INPUTS 3
typedef struct PccMlpInput {
    const void *ptr;
    int dtype;
    int channels;
    int64_t ld;
    int row_div;
} PccMlpInput;
int64_t pcc_mlp_packed_bytes(int cin, int cout);
int pcc_mlp_pack_weights_f32(const float *w, const float *bias, int cin, int cout, void *packed, void *stream);
int pcc_mlp_chain_f32(const float *x, int64_t rows, int ldx, const PccMlpLayer *layers, int n_layers, int group,
                      float *out, void *stream);
/* General form: several input segments, output fp32 (out_dtype 0) or bf16 (out_dtype 1). */
int pcc_mlp_chain(const PccMlpInput *inputs, int n_inputs, int64_t rows, const PccMlpLayer *layers, int n_layers,
                  int group, void *out, int out_dtype, void *stream);

/*
 * Diagnostics (tools/time_chain.py): when given a device buffer of 256 int64, CTA 0 / thread 0 of the next
 * pcc_mlp_chain launches stores clock64() at its phase boundaries; pass NULL to switch it off (the default).
 */
void pcc_debug_mlp_timing(long long *device_buf);
/* same for the warp-specialised SetAbstraction chain (chain_ws.cu): 1024 int64, MMA warp ticks at [512..) */
void pcc_debug_ws_timing(long long *device_buf);

/*
 * Fused tail of pn_kit.PointNet (/root/reference/pn_kit.py:136-143 as configured by AE.py:17): per position
 *     y = W3 . relu(W2 . x + b2) + b3   (256 -> 512 -> cout <= 16),   then max over every run of 256 positions.
 * x [rows, 256] bf16 (row pitch ldx elements; rows % 256 == 0), w2 [512, 256] bf16 row-major, b2 [512] fp32,
 * w3_packed = pcc_mlp_pack_weights_f32(cin = 512, cout), out [rows / 256, cout] fp32.  The 512-wide activation stays on
 * the SM (TMEM / shared memory); W2 is streamed from L2 by TMA.
 */
int pcc_pn_tail_bf16(const void *x, int64_t rows, int64_t ldx, const void *w2_bf16, const float *b2, const void *w3_packed,
                     int cout, int relu3, float *out, void *stream);

/*
 * Per-cloud eval.py metrics from the by-products of pcc_chamfer_fwd_f32(decompressed, original): out [B,3] float64 =
 * (Chamfer on the (p - min) / (max - min) normalised clouds, eval.py:199-205; D1 PSNR in dB, eval.py:68-92; D1 MSE).
 * dx [B,P1] = recon -> original squared distances, per_cloud [B], bbox [B,6] = (min xyz, max xyz) of the original.
 */
int pcc_eval_metrics_f32(const float *dx, const float *per_cloud, const float *bbox, int B, int P1, double *out, void *stream);

/*
 * Octree coding of the patch centres, the stage between FPS and kNN patching.
 * pcc_octree_encode_f32 replaces octree_np.encode (/root/reference/octree_np.py:10-45, with its getDecodeFromPc snap,
 * :114-133) and the per-cloud depth search of pn_kit.encode_sampled_np (/root/reference/pn_kit.py:380-401; call sites
 * train.py:176, compress.py:98) at scale = 1:  fixed_depth = 0 searches depth 1..16 for the first depth whose stream has
 * nbits / n_points > min_bpp with all S centres in distinct cells (depth 16 if none does); fixed_depth in 1..16 encodes at
 * that depth.  centres [B, S, 3] must lie in [0, 1) (octree_np.py:5-7); a cloud that does not gets out_depth = -1,
 * out_nbits = 0.  S <= 8192.
 *   out_bits  [B, max_bits] uint8, one byte per bit (the reference's np.uint8 array), zero beyond out_nbits[b];
 *             max_bits >= pcc_octree_max_bits(S) (or 1 + 8 * fixed_depth * S);
 *   out_nbits [B] int32, out_depth [B] int32 (depth of the returned code);
 *   out_bytes (nullable) [B, (max_bits + 7) / 8]: pn_kit.binary_array_to_byte_array of the stream (pn_kit.py:463-467;
 *             compress.py:144), (out_nbits + 7) / 8 bytes used;
 *   out_quant (nullable) [B, S, 3]: octree_np.getDecodeFromPc of every centre at the chosen depth, input order;
 *   out_rec_ref (nullable) [B, 64, 3]: what the reference's decoder returns for this stream (see below);
 *   out_stream_xyz (nullable) [B, S, 3]: what pcc_octree_decode_f32 mode 1 returns for this stream (the distinct leaf
 *             centres in stream order, the last one repeated up to S rows) -- the centres a decoder will see.
 * pcc_octree_decode_f32: mode 0 = octree_np.decode exactly as written (/root/reference/octree_np.py:47-112;
 * pn_kit.decode_sampled_np, pn_kit.py:424-431): it consumes the first 8 bits only, returns depth-1 octant centres and pads
 * to 64 rows (cap must be 64); mode 1 = the inverse of the encoder (leaf centres in stream order, the last one
 * repeated after out_count[b], at most cap).  bits [B, max_bits] uint8, nbits [B] int32, out_xyz [B, cap, 3].
 */
int pcc_octree_max_bits(int S);
int pcc_octree_encode_f32(const float *centres, int B, int S, int n_points, double min_bpp, int fixed_depth,
                          uint8_t *out_bits, int max_bits, int32_t *out_nbits, int32_t *out_depth, uint8_t *out_bytes,
                          float *out_quant, float *out_rec_ref, float *out_stream_xyz, void *stream);
int pcc_octree_decode_f32(const uint8_t *bits, const int32_t *nbits, int B, int max_bits, int mode, int cap, float *out_xyz,
                          int32_t *out_count, int32_t *out_depth, void *stream);

/*
 * One wide shared-MLP layer as a streamed tensor-core GEMM (csrc/gemm_ws.cu): out = act(a . w^T + bias).
 * Replaces the Conv2d(1x1)+BatchNorm(folded)+ReLU layers of PointnetSAModule (/root/reference/pointnet_sa_module.py:39-56,87-91
 * as configured by PPPF_AE.py:29-33) and the Linear layers of AE.inv_pool (/root/reference/AE.py:19-26) whose weights do not
 * fit in shared memory beside the activations.
 * a [M, K] bf16 (row pitch lda), w [N, K] bf16 (row pitch ldw), bias [N] fp32; K % 64 == 0, N % 128 == 0 (pad with zero
 * columns / use pcc_gather_concat_bf16); pitches multiples of 8 elements, 16-byte aligned bases.
 * group <= 0: out [M, N] bf16 (row pitch ld_out).  group == 1: out [M, N] fp32, dense (logits / per-cloud vectors that must
 * not be rounded to bf16).  group > 1: the max over every run of `group` consecutive rows
 * (pointnet_sa_module.py:91 torch.max over nsample): out [M / group, N] fp32, group % 32 == 0 and either a divisor or a
 * multiple of 128, M % group == 0, relu required when group > 32.
 * Numerics: operands bf16, products accumulated in fp32 (tolerance stated in tests/test_gpu_mlp.py).
 */
int pcc_linear_bf16(const void *a, int64_t M, int K, int64_t lda, const void *w, int64_t ldw, const float *bias, int N, int relu,
                    int group, void *out, int64_t ld_out, void *stream);

/*
 * Skinny fp32 Linear for per-cloud vectors (csrc/small_ops.cu): out[M, N] = act(x[M, K] . w[N, K]^T + bias) with M = clouds.
 * Replaces nn.Linear enc_proj / dec_proj of PPPF_AE (/root/reference/PPPF_AE.py:122-123,139,145), the latent columns of
 * FoldingNet's first Conv1d of each stage (PPPF_AE.py:58-59,68-69,100-109: the latent is identical for every grid point) and
 * the pooled-feature columns of AE.ConditionalProbabilityModel.model_mlp[0] (/root/reference/AE.py:99,115-116).
 * Weight-bandwidth bound, fp32 FMA chains in k order: a row's result is independent of M (batch invariant).  bias nullable.
 */
int pcc_linear_small_f32(const float *x, int M, int K, int64_t ldx, const float *w, int64_t ldw, const float *bias, int N, int relu,
                         float *out, int64_t ld_out, void *stream);

/*
 * First layer of a stage whose input is cat([n_local per-point values, a per-cloud vector tiled over the n_pts points of the
 * cloud]) (PPPF_AE.py:100-101,106-107 torch.cat + Conv1d; AE.py:115-116): out[r, c] = act(per_cloud[r / n_pts, c] +
 * sum_j local[r, j] * w[c, j]) in fp32, written as bf16 rows (pitch ld_out, columns C..ld_out zero) -- the A operand of
 * pcc_linear_bf16 / pcc_mlp_chain.  local [M, n_local] fp32 (pitch ld_local, n_local <= 4), w [C, n_local] fp32 (pitch ldw: a
 * column slice of the layer's weight), per_cloud [M / n_pts, C] fp32 (bias included, from pcc_linear_small_f32).
 */
int pcc_fold_first_bf16(const float *local, int n_local, int64_t ld_local, const float *w, int64_t ldw, const float *per_cloud,
                        int64_t M, int n_pts, int C, int relu, void *out, int64_t ld_out, void *stream);

/*
 * pcc_knn_f32 for scene-scale candidate clouds (compress.py:70-74 on a Stanford3D / S3DIS scene, pppe_pcd_ae.py:599 on a whole
 * cloud): identical outputs, bit for bit, but the candidates are visited through a 32^3 uniform grid built over p (cell, then
 * shell after shell, until the K-th distance is inside the scanned block) -- ~10^3 distance evaluations per query instead of P2.
 * workspace: pcc_knn_grid_workspace_bytes(B, P2) bytes, 16-byte aligned, caller-owned.
 */
int64_t pcc_knn_grid_workspace_bytes(int B, int P2);
int pcc_knn_grid_f32(const float *q, const float *p, int B, int P1, int P2, int K, float *out_d2, int64_t *out_idx, float *out_nn,
                     int centre_sub, float nn_scale, void *workspace, void *stream);

/*
 * pn_kit.PointNet of the AE in one launch (/root/reference/pn_kit.py:124-144 as AE.py:39 calls it on cat((xyz, feat))):
 * out[patch, cout] = max over the patch's 256 positions of W3 . relu(W2 . relu(W1 . relu(W0 . [feat | xyz] + b0) + b1) + b2) + b3,
 * widths 131 -> 128 -> 256 -> 512 -> cout <= 16.  feat [rows, 128] bf16 (the SetAbstraction output), xyz [rows, 3] fp32,
 * rows % 256 == 0.  w0f [128, 128] bf16 = the feature columns of the first layer, w0_packed = pcc_mlp_pack_weights_f32 of the
 * rotated layer [feat | xyz] (cin = 131: xyz columns and bias), w1 [256, 128] / w2 [512, 256] bf16 row-major, b1 / b2 fp32,
 * w3_packed = pcc_mlp_pack_weights_f32(cin = 512).  No intermediate activation reaches HBM (the two-launch route
 * pcc_mlp_chain + pcc_pn_tail_bf16 writes and re-reads a [rows, 256] bf16 tensor).
 */
int pcc_pointnet_fused_bf16(const void *feat, int64_t rows, int64_t ld_feat, const float *xyz, int64_t ld_xyz, const void *w0f_bf16,
                            const void *w0_packed, const void *w1_bf16, const float *b1, const void *w2_bf16, const float *b2,
                            const void *w3_packed, int cout, int relu3, float *out, void *stream);

/*
 * Weight gradient of one shared-MLP layer (the backward of the Conv2d(1x1) / Linear layers the reference trains under autograd,
 * /root/reference/train.py:193-221): c[Na, Nb] = a[M, Na]^T . b[M, Nb] (fp32, overwritten; row pitch ldc) and, when colsum is
 * given, colsum[Na] = column sums of a (the bias gradient).  a = the output gradient, b = the layer's input, both bf16
 * row-major with widths / pitches that are multiples of 8 elements, 16-byte aligned.  The contraction over the M rows runs on
 * the tensor cores straight from the row-major tensors (MN-major operands); long M with a small Na x Nb is split over the SMs
 * and reduced with fp32 atomics (the sum order is then not fixed: results are reproducible to fp32 rounding, not bit-exact).
 */
int pcc_wgrad_bf16(const void *a, int64_t lda, int Na, const void *b, int64_t ldb, int Nb, int64_t M, float *c, int64_t ldc,
                   float *colsum, void *stream);

/*
 * Training forms of the streamed layer and of the pooling (the reference trains these layers under autograd,
 * /root/reference/train.py:193-221; here every contraction of the forward AND backward pass is one of these kernels).
 * pcc_linear_train_bf16: pcc_linear_bf16 with a bf16 [M, n_store] result (n_store % 8 == 0, <= N: columns past it are not
 *   written -- narrow layers inside the 128-column granule) and an optional mask [M, >= n_store] bf16: out = 0 where mask <= 0,
 *   i.e. the ReLU backward of the layer below fused into the data-gradient GEMM  dX = (dY . W) * [X > 0].
 * pcc_groupmax_fwd_bf16: pooled[M / group, C] fp32 = max over each run of `group` rows of x [M, C] bf16 (row pitch ldx), arg =
 *   the first row of the run that attains it (pn_kit.py:139-143, 207).  pcc_groupmax_bwd_bf16: dy [M, ld_dy] bf16 = dout routed
 *   to the arg-max rows (and only where pooled > 0 when `pooled` is given: a ReLU preceded the max); columns C .. ld_dy are zero.
 */
int pcc_linear_train_bf16(const void *a, int64_t M, int K, int64_t lda, const void *w, int64_t ldw, const float *bias, int N, int relu,
                          void *out, int64_t ld_out, int n_store, const void *mask, int64_t ld_mask, void *stream);
int pcc_groupmax_fwd_bf16(const void *x, int64_t M, int C, int64_t ldx, int group, float *pooled, int16_t *arg, void *stream);
int pcc_groupmax_bwd_bf16(const float *dout, const float *pooled, const int16_t *arg, int64_t M, int C, int group, void *dy,
                          int64_t ld_dy, void *stream);

/*
 * pn_kit.SetAbstraction with S == N (/root/reference/pn_kit.py:181-207 as AE.sa calls it, AE.py:38) in two launches that
 * never materialise the grouped tensor: the in-patch kNN table as bytes, then the shared MLP gathering from the patch itself.
 * pcc_knn_patch_u8: patches [BS, P, 3] fp32, K in {8, 16}, K <= P <= 256 -> out_idx [BS, P, K] uint8: the K nearest points of
 *   every point inside its own patch, (d2, idx) order, bit-exact to pytorch3d.ops.knn_points (pn_kit.py:190).
 * pcc_sa_chain_indexed: position (point g, neighbour n) = patches[patch(g) * P + idx8[g, n]] - patches[g] (pn_kit.py:191, fp32
 *   subtraction) -> 3 -> 32 -> 64 -> 128 (ReLU) -> max over the 16 neighbours (pn_kit.py:196-207); layers as for pcc_mlp_chain
 *   (layer 0 needs w_f32 / b_f32).  out [points, 128] fp32 (out_dtype 0) or bf16 (1).  Any other shape: PCC_ERR_UNSUPPORTED
 *   (the general route is pcc_knn_f32 with out_nn + pcc_mlp_chain).  Result bit-identical to that route.
 */
int pcc_knn_patch_u8(const float *patches, int BS, int P, int K, uint8_t *out_idx, void *stream);
int pcc_sa_chain_indexed(const float *patches, const uint8_t *idx8, int64_t points, int pts_per_patch, const PccMlpLayer *layers,
                         int n_layers, void *out, int out_dtype, void *stream);

/*
 * Backward pass of pcc_sa_chain_indexed with respect to the six parameter tensors -- what autograd computes through
 * pn_kit.SetAbstraction's three Conv2d(1x1) + ReLU layers and the max over the neighbours when the reference trains
 * (/root/reference/train.py:193-221 -> pn_kit.py:196-207) -- in one kernel that recomputes the activations tile by tile
 * (csrc/sa_bwd.cu) instead of storing them in the forward pass.  patches / idx8 / points / pts_per_patch as in the forward call;
 * w0 [32, 3], b0 [32], w1 [64, 32], b1 [64], w2 [128, 64], b2 [128] fp32 (the master weights; operands are rounded to bf16 as in
 * the forward kernel); grad_out = dLoss / d(out): `points` rows of 128 values, grad_ld elements apart (>= 128), fp32 (grad_dtype 0)
 * or bf16 (grad_dtype 1: the slice of the next stack's input gradient autograd hands back, read in place).  The gradients are ADDED
 * into dw0 .. db2 (fp32, same shapes): zero them first.  The grouped coordinates are data, so there is no input gradient.
 */
int pcc_sa_chain_indexed_bwd(const float *patches, const uint8_t *idx8, int64_t points, int pts_per_patch, const float *w0,
                             const float *b0, const float *w1, const float *b1, const float *w2, const float *b2,
                             const void *grad_out, int grad_dtype, int64_t grad_ld, float *dw0, float *db0, float *dw1, float *db1,
                             float *dw2, float *db2, void *stream);

/*
 * Grouping of one PointNet++ set-abstraction level (/root/reference/pointnet_sa_module.py:73-85: group_points of the features
 * and of xyz, torch.cat): out[b * M + j, :] = [feat[b, idx[b, j], 0..C) | xyz[b, idx[b, j], 0..3) | 0 ...] as bf16 rows of
 * kpad columns (kpad % 8 == 0) -- the A operand of pcc_linear_bf16.  feat [B, N, C] fp32 or NULL (C = 0), xyz [B, N, 3] fp32 or
 * NULL, idx [B, M] int64; negative indices read point 0 (pointnet_sa_module.py:27).
 * centre (nullable) [B, M / nsample, 3] fp32: the query every run of `nsample` rows was grouped around; when given, the xyz
 * columns hold fl(xyz[idx] - centre) -- the recentred grouping of pppe_pcd_ae.PointNetSetAbstraction
 * (/root/reference/pppe_pcd_ae.py:599-607: knn_points(return_nn) - new_xyz, index_points of the features, torch.cat).
 */
int pcc_gather_concat_bf16(const float *feat, int C, const float *xyz, const int64_t *idx, int B, int N, int64_t M, int kpad,
                           void *out, const float *centre, int nsample, void *stream);

/*
 * The remaining eval.py metrics, computed from outputs of pcc_knn_f32 / pcc_chamfer_fwd_f32.
 * pcc_normals_pca_f32: Open3D estimate_normals(KDTreeSearchParamKNN(knn = 30)) of /root/reference/eval.py:58-59 --
 *   nn [rows, knn, 3] = the knn nearest points of every point of the original cloud, itself included (pcc_knn_f32's out_nn
 *   without recentring); out_normals [rows, 3] = unit eigenvector of the smallest eigenvalue of their covariance (double
 *   arithmetic; sign arbitrary, eval.py squares the projection).
 * pcc_p2plane_f32: eval.py:68-92 -- ix [B, P1] = nearest original point of every reconstructed point (pcc_chamfer_fwd_f32's
 *   out_ix for (recon, orig)); out [B, 2] float64 = (mean (diff . normal)^2, 10 log10(|bbox diag|^2 / mse)); bbox [B, 6].
 * pcc_uc_f32: the variance ratio that ends calc_uc (eval.py:127-151) -- d2_in / d2_dec [B, n] = squared distance of every point
 *   of the 1024-point region around point 0 to its nearest other region point (pcc_knn_f32 with K = 2, column 1), for the
 *   input and the decompressed cloud; out [B] float64 = var(sqrt(d2_dec)) / var(sqrt(d2_in)), population variance (np.var).
 */
int pcc_normals_pca_f32(const float *nn, int64_t rows, int knn, float *out_normals, void *stream);
int pcc_p2plane_f32(const float *recon, const float *orig, const int64_t *ix, const float *normals, const float *bbox, int B,
                    int P1, int P2, double *out, void *stream);
int pcc_uc_f32(const float *d2_in, const float *d2_dec, int B, int n, double *out, void *stream);

/*
 * Entropy stage: the host code between the probability model and the .p.bin file.
 * pcc_pmf_to_cdf_u16: pn_kit.pmf_to_cdf (/root/reference/pn_kit.py:452-461; compress.py:134, decompress.py:92) fused with
 *   torchac's float -> 16-bit CDF conversion (encode_float_cdf(..., needs_normalization = True)): pmf [rows, L] fp32 ->
 *   cdf [rows, L + 1] uint16.  pcc_cdf_to_u16 converts a float CDF [rows, Lp] the caller already holds.
 * pcc_range_encode_u16 / pcc_range_decode_u16: torchac.encode_float_cdf / decode_float_cdf's coder (compress.py:136,
 *   decompress.py:93), one byte stream per cloud: cdf [B, n_sym, Lp] uint16, sym [B, n_sym] int16 in [0, Lp - 2],
 *   bytes [B, cap] with cap >= 2 * n_sym + 8, nbytes [B] int32.
 * torchac is a third-party package absent from /root/reference: its published algorithm is restated (parity unpinned).
 */
int pcc_pmf_to_cdf_u16(const float *pmf, int64_t rows, int L, uint16_t *out_cdf, void *stream);
int pcc_cdf_to_u16(const float *cdf_float, int64_t rows, int Lp, uint16_t *out_cdf, void *stream);
int pcc_range_encode_u16(const uint16_t *cdf, const int16_t *sym, int B, int n_sym, int Lp, uint8_t *out_bytes, int cap,
                         int32_t *out_nbytes, void *stream);
int pcc_range_decode_u16(const uint16_t *cdf, const uint8_t *bytes, const int32_t *nbytes, int B, int n_sym, int Lp, int cap,
                         int16_t *out_sym, void *stream);

/*
 * The quantiser between encoder and decoder (/root/reference/AE.py:42-45, compress.py:125-127):
 * latent = sigmoid(raw) * spread - spread / 2 with spread = L - 0.2, latent_q = round(latent); raw [rows, d] fp32.
 * out_q_bf16 (nullable) [rows, kpad]: the rounded latent as bf16 rows zero padded to kpad columns (operand of pcc_linear_bf16).
 */
int pcc_quantise_latent_f32(const float *raw, int64_t rows, int d, int kpad, float spread, float *out_latent, float *out_q,
                            void *out_q_bf16, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* PCC_B200_H */
