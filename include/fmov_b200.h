/* fmov_b200 — C ABI of the B200-native NeuS train-step hot path.
 *
 * Drop-in boundary for the path BASELINE.json `north_star` names.  The reference
 * (diegointel/fmov_pose) has NO native/FFI layer: its seam is the Python interface of
 * models/renderer.py, models/fields.py, models/dataset.py and the pose modules.
 * The Python host code in fmov_pose_b200/models/ keeps those interfaces and calls the entry
 * points below through ctypes; each entry point cites the reference lines it replaces.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the comment says "host";
 *   - `stream` is a cudaStream_t passed as void*;
 *   - return value: 0 = ok, <0 = error (fmov_last_error() gives the message); nothing throws,
 *     nothing allocates, nothing is retained after the call returns (stream-ordered);
 *   - fp32 tensors are dense row-major; "image" tensors are opaque fp16/bf16 tile images
 *     (fmov_pose_b200/csrc/fmov_common.cuh) produced and consumed only by this library.
 */
#ifndef FMOV_B200_H_
#define FMOV_B200_H_

#ifdef __cplusplus
extern "C" {
#endif

/* ---- library ------------------------------------------------------------------------- */
const char* fmov_last_error(void);
int fmov_version(void);
/* number of kernels this library has launched in this process (a launch recorded into a CUDA graph counts once) */
unsigned long long fmov_launch_count(void);

/* ---- weight / tile images (host glue for the MLP kernels) ----------------------------- */
/* fp32 matrix -> no-swizzle operand image [npad x 64*kblocks]; see api.cu. seg_* are HOST arrays.
 * bf16: 0 = fp16 image, 1 = bf16 image, 2 = fp16 image of the residual x - fp16(x) (split-precision chains). */
int fmov_pack_image(const float* src, long long stride_n, long long stride_k, int n_valid, int row_off, int nseg,
                    const int* seg_dst, const int* seg_src, const int* seg_len, float scale, int bf16, void* dst,
                    int npad, int kblocks, void* stream);
int fmov_ti_from_rowmajor(const float* src, long long P, int cols, int ld, int kblocks, int bf16, void* dst, void* stream);
int fmov_ti_to_rowmajor(const void* src, long long P, int cols, int ld, int kblocks, int bf16, float* dst, void* stream);
/* tcgen05 descriptor self-test (tests only). */
int fmov_selftest_gemm(const void* a_img, const void* b_img, int n, int kblocks, int a_bf16, int b_bf16, int mode,
                       float* out, void* stream);

/* ---- weight normalisation of all layers in one launch ------------------------------------------------- */
/* nn.utils.weight_norm(dim=0) of every linear layer (models/fields.py:81-82, 160-161): W = g * v / ||v|| per output row,
 * and its backward dL/dW -> (dL/dv, dL/dg).  Every pointer argument is a HOST array of n_layers (<= 16) device pointers;
 * layer i: v [rows[i], cols[i]], g [rows[i]] (weight_g), W / dW / dv like v, norm / dg [rows[i]].                     */
int fmov_weight_norm_fwd(int n_layers, const float* const* v, const float* const* g, const int* rows, const int* cols,
                         float* const* W, float* const* norm, void* stream);
int fmov_weight_norm_bwd(int n_layers, const float* const* v, const float* const* g, const int* rows, const int* cols,
                         const float* const* norm, const float* const* dW, float* const* dv, float* const* dg,
                         void* stream);

/* ---- SDF value query: SDFNetwork.sdf under no_grad ------------------------------------- */
/* replaces models/fields.py:88-107 at the call sites models/renderer.py:424-428 (coarse),
 * :230-232 (up-sample rounds) and :506 via extract_fields :9-37 (dense grid).
 * wblob: forward weight images of lin0..lin7 (fmov_sdf_fwd_blob_bytes / _offset give the layout);
 * bias8x256: biases of lin0..lin7 zero-padded to 256; w8_row0: row 0 of lin8; b8: lin8.bias (device, [0] used).             */
long long fmov_sdf_fwd_blob_bytes(void);
long long fmov_sdf_fwd_blob_offset(int layer);
int fmov_sdf_query_points(const float* pts, long long P, const void* wblob, const float* bias8x256, const float* w8_row0,
                          const float* b8, float in_scale, float out_scale, float* out, void* stream);
int fmov_sdf_query_rays(const float* rays_o, const float* rays_d, const float* z, long long B, int S, int z_stride,
                        int z_off, const void* wblob, const float* bias8x256, const float* w8_row0, const float* b8,
                        float in_scale, float out_scale, float* out, void* stream);
/* bmin3/bmax3 are HOST float[3]; points first..first+count of the x-major res^3 grid. */
/* fmov_sdf_query_rays on the CTA-pair engine (clusters of two CTAs, tcgen05.mma.cta_group::2): the hierarchical-sampling
 * queries of the train step.  wblob_pair = the FP0..FP7 region of the fine-stage blob (fmov_sdf_pair_blob_offset(),
 * fmov_sdf_pair_blob_bytes() bytes; 0 bytes when the library was built without the pair engine): half-major images, biases in their bias slices. */
long long fmov_sdf_pair_blob_bytes(void);
long long fmov_sdf_pair_blob_offset(void);   /* where that region starts inside the fmov_pack_all blob */
int fmov_sdf_query_rays_pair(const float* rays_o, const float* rays_d, const float* z, long long B, int S, int z_stride,
                             int z_off, const void* wblob_pair, const float* w8_row0, const float* b8, float in_scale,
                             float out_scale, float* out, void* stream);
int fmov_sdf_query_grid(const float* bmin3, const float* bmax3, int res, long long first, long long count,
                        const void* wblob, const float* bias8x256, const float* w8_row0, const float* b8, float in_scale,
                        float out_scale, float* out, void* stream);

/* split-precision value chains for SDFNetwork.sdf / extract_fields (models/fields.py:106-107, models/renderer.py:9-37, :506),
 * measured against the reference's own 40^3 grid on the +-1.01 box (north_star: SDF <= 1e-3; profiles/r2b_grid_modes.txt):
 *   wblob_lo != NULL  full split: every layer accumulates hi*W_hi + lo*W_hi + hi*W_lo with hi = fp16(v), lo = fp16(v - hi):
 *                     1.3e-5, 0.30 G queries/s per GPU.  wblob_lo = residual images W - fp16(W) in wblob's layout
 *                     (fmov_pack_image with bf16 = 2)
 *   wblob_lo == NULL  activation split: (hi + lo)*W with plain fp16 weights, one pass over the weight stream: 3.6e-4,
 *                     0.40 G queries/s (the default of the Python mirror's sdf() / extract_fields)
 * (the plain chain, fmov_sdf_query_points / _grid: 6.7e-4 on that box, 1e-3 at |x| ~ 2, 0.74 G queries/s) */
int fmov_sdf_query_points_precise(const float* pts, long long P, const void* wblob, const void* wblob_lo,
                                  const float* bias8x256, const float* w8_row0, const float* b8, float in_scale,
                                  float out_scale, float* out, void* stream);
int fmov_sdf_query_grid_precise(const float* bmin3, const float* bmax3, int res, long long first, long long count,
                                const void* wblob, const void* wblob_lo, const float* bias8x256, const float* w8_row0,
                                const float* b8, float in_scale, float out_scale, float* out, void* stream);

/* ---- pose + ray generation ------------------------------------------------------------ */
/* mode 1: LearnPoseGF tail  c2w = [Exp(rot)|trans] @ [R0 | scale*t0]  (models/picture_pose.py:176-186,
 *         models/batch_lie_group_helper.py:19-47); init34 = rows 0..2 of init_c2w[cam] (row stride 4)
 * mode 2: BARF  c2w = compose_pair(se3_to_SE3(se3), noise_pose)  (models/camera.py:89-102, 53-60;
 *         exp_runner.py:419-424); init34 = noise pose rows (row stride 4)                           */
int fmov_pose_fwd(int mode, const float* rot, const float* trans, const float* scale, const float* init34,
                  const float* se3, float* c2w34, void* stream);
int fmov_pose_bwd(int mode, const float* rot, const float* trans, const float* scale, const float* init34,
                  const float* se3, const float* g_c2w34, float* g_rot, float* g_trans, float* g_scale, float* g_se3,
                  void* stream);
/* LearnPoseGF.forward (models/picture_pose.py:140-186) in ONE launch per direction: Gaussian-Fourier features of the frame
 * index cid[0] (device int64) -> lin1 (256->64) GELU -> lin2 (64->64) GELU -> heads -> Rodrigues tail (mode 1 above) with
 * init_all[cid] ([N,4,4]; NULL = identity).  Heads (HOST arrays of n_heads device pointers + row counts), rows concatenated
 * as rot(3) trans(3) [scale(1)]: {lin3 (6)} or {lin3_rot (3), lin3_trans (3), lin3_scale (1)}; rot rows are multiplied by
 * rot_k (pi, or pi/6 with small_rot).  save: fmov_pose_gf_save_floats() floats written by fwd, read by bwd.  Gradient
 * pointers may be NULL (frozen parameters); dWh/dbh are HOST arrays of device pointers (entries may be NULL).       */
int fmov_pose_gf_save_floats(void);
int fmov_pose_gf_fwd(const long long* cid, const float* b, const float* W1, const float* b1, const float* W2,
                     const float* b2, int n_heads, const float* const* Wh, const float* const* bh, const int* rows,
                     float rot_k, const float* init_all, float* save, float* c2w34, void* stream);
int fmov_pose_gf_bwd(const long long* cid, const float* b, const float* W1, const float* b1, const float* W2,
                     const float* b2, int n_heads, const float* const* Wh, const float* const* bh, const int* rows,
                     float rot_k, const float* init_all, const float* save, const float* g_c2w34, float* dW1, float* db1,
                     float* dW2, float* db2, float* const* dWh, float* const* dbh, void* stream);
/* Dataset.gen_random_rays_at ray math (models/dataset.py:656-671) + near_far_from_sphere (:835-842), with the
 * pose evaluated in-kernel (mode 0: c2w34 given). px/py are int64 pixel coordinates.                 */
int fmov_raygen_fwd(int mode, const float* c2w34, const float* rot, const float* trans, const float* scale,
                    const float* init34, const float* se3, const float* intr_inv, int intr_stride, const long long* px,
                    const long long* py, long long B, float* rays_o, float* rays_d, float* near, float* far,
                    float* c2w_out, void* stream);
int fmov_raygen_bwd(const float* intr_inv, int intr_stride, const long long* px, const long long* py, long long B,
                    const float* rays_o, const float* rays_d, const float* g_o, const float* g_d, const float* g_near,
                    const float* g_far, float* g_c2w34, void* stream);

/* same with float (sub-pixel) pixel coordinates: the LoFTR flow matches of Dataset.gen_random_ray_pairs_at
 * (models/dataset.py:712-715, 728-760)                                                              */
int fmov_raygen_xy_fwd(int mode, const float* c2w34, const float* rot, const float* trans, const float* scale,
                       const float* init34, const float* se3, const float* intr_inv, int intr_stride, const float* px,
                       const float* py, long long B, float* rays_o, float* rays_d, float* near, float* far,
                       float* c2w_out, void* stream);
int fmov_raygen_xy_bwd(const float* intr_inv, int intr_stride, const float* px, const float* py, long long B,
                       const float* rays_o, const float* rays_d, const float* g_o, const float* g_d, const float* g_near,
                       const float* g_far, float* g_c2w34, void* stream);

/* ---- hierarchical sampling --------------------------------------------------------------- */
/* coarse z + per-ray jitter (models/renderer.py:385-405); t_rand = the raw U[0,1) draw or NULL  */
int fmov_sample_coarse(const float* near, const float* far, const float* t_rand, long long B, int n_samples,
                       int z_stride, float* z, void* stream);
/* one importance round: merge z[:, :n_sorted] with the sorted tail z[:, n_sorted:n_sorted+n_tail] (cat_z_vals,
 * models/renderer.py:222-242; sdf permuted alongside when with_sdf), then up_sample + sample_pdf(det)
 * (:168-220, :54-86) writing n_new samples after the merged ones.                                  */
int fmov_sample_round(const float* rays_o, const float* rays_d, float* z, float* sdf, long long B, int z_stride,
                      int n_sorted, int n_tail, int with_sdf, int n_new, float inv_s, void* stream);
/* sample_pdf(bins [B,n], weights [B,n-1], n_new, det=True) -> out [B,n_new]: models/renderer.py:54-86 on caller-given
 * weights (the inverse-CDF stage of fmov_sample_round, exposed for known-answer tests). */
int fmov_sample_pdf(const float* bins, const float* weights, long long B, int n, int n_new, float* out, void* stream);

/* ---- compositing + losses ------------------------------------------------------------------ */
/* render_core tail (models/renderer.py:261-272, 290-358) + render() reductions (:477-498).
 * inv_s: device scalar, already clipped to [1e-6,1e6]; bg: device float[3] or NULL;
 * eik_partial [B,2]: per-ray (sum relax*(|n|-1)^2, sum relax).                                      */
int fmov_composite_fwd(long long B, int S, const float* rays_o, const float* rays_d, const float* z, const float* sdf,
                       const float* nrm, const float* rgb, const float* inv_s, float sample_dist, float cos_anneal,
                       const float* bg, float* color, float* weight_sum, float* weight_max, float* depth,
                       float* weights, float* cdf, float* inside, float* mid_z, float* pts, float* eik_partial,
                       void* stream);
int fmov_composite_bwd(long long B, int S, const float* rays_o, const float* rays_d, const float* z, const float* sdf,
                       const float* nrm, const float* rgb, const float* inv_s, float sample_dist, float cos_anneal,
                       const float* bg, const float* g_color, const float* g_wsum, const float* g_depth,
                       const float* g_weights, const float* g_eik, const float* eik_den, const float* g_nrm_ext,
                       float* d_sdf, float* d_nrm, float* d_rgb, float* d_dir, float* d_dist, float* d_mid,
                       float* d_invs, void* stream);
/* exp_runner.py:562-599: per-ray masked-L1 and BCE terms + gradients for loss = colour + mask_weight*bce */
int fmov_loss_fwd_bwd(const float* color, const float* weight_sum, const float* true_rgb, const float* mask, long long B,
                      const float* mask_sum, long long n_rays_global, float mask_weight, float* partial, float* g_color,
                      float* g_wsum, void* stream);
/* autograd of pts = o + d*mid_z (models/renderer.py:269-272) reduced per ray */
int fmov_ray_reduce_bwd(const float* d_pts, const float* d_dirs, const float* d_dir_tc, const float* d_dist,
                        const float* d_mid, const float* rays_d, const float* z, long long B, int S, float sample_dist,
                        float* d_o, float* d_d, float* d_z, void* stream);

/* ---- flow / reprojection and unit-sphere losses ------------------------------------ */
/* exp_runner.py:605-688: err[r] = sum_j weights[r,j] * (pi(K (R_w p_rj + t_w)) - xy[r]),  p_rj = o_r + d_r*mid_z_rj
 * (mid_z from z and sample_dist as models/renderer.py:261-267).  w2c34 [3,4] = rows 0..2 of inverse(c2w) of the matched
 * frame (device), K33 = its intrinsics (device, row stride k_stride), xy [B,2] the matched pixels.  The caller applies
 * F.l1_loss(err, 0) * flow_weight.  backward: d_weights [B,S] (NULL: detach_flow_on_sdf), d_o/d_d [B,3], d_z [B,S]
 * (NULL unless z carries gradient, i.e. n_importance == 0), d_w2c34 [12] (zeroed by the call, summed over rays).  */
int fmov_flow_fwd(long long B, int S, const float* rays_o, const float* rays_d, const float* z, float sample_dist,
                  const float* weights, const float* w2c34, const float* K33, int k_stride, const float* xy, float* err,
                  void* stream);
int fmov_flow_bwd(long long B, int S, const float* rays_o, const float* rays_d, const float* z, float sample_dist,
                  const float* weights, const float* w2c34, const float* K33, int k_stride, const float* xy,
                  const float* g_err, float* d_weights, float* d_o, float* d_d, float* d_z, float* d_w2c34, void* stream);
/* exp_runner.py:714-724: partial2 = (sum |w| over samples with |pts| > 1, their count) (zeroed by the call; NULL to
 * skip); d_weights = g_scale[0] * sign(w) on those samples, 0 elsewhere (NULL to skip; g_scale device scalar).    */
int fmov_unit_sphere_fwd_bwd(long long B, int S, const float* rays_o, const float* rays_d, const float* z,
                             float sample_dist, const float* weights, const float* g_scale, float* partial2,
                             float* d_weights, void* stream);

/* ---- fine stage: SDF value + feature + analytic normal + colour MLP, forward and backward -------- */
/* forward: replaces sdf_network(pts), sdf_network.gradient(pts), color_network(...) of render_core
 * (models/renderer.py:277-288; models/fields.py:88-124, 166-193).  backward + fmov_dw: the matching part of
 * loss.backward() (exp_runner.py:802).  Points are o + d*mid_z of the [B,S] samples (renderer.py:261-272).
 * wblob: weight images (fmov_fine_image_info gives id -> offset/npad/kblocks; packed with fmov_pack_image);
 * stash: HOST array of fmov_fine_stash_count() device pointers, tensor i holding
 *        ceil(B*S/128) * fmov_fine_stash_blocks(i) * 16384 bytes (activation / gradient tile images).       */
int fmov_fine_image_count(void);
int fmov_fine_image_info(int id, long long* offset, int* npad, int* kblocks);
long long fmov_fine_blob_bytes(void);
/* one launch: every operand image + the fp32 side arrays from the 28 effective-weight tensors (HOST array of device
 * pointers W_sdf[9], b_sdf[9], W_col[5], b_col[5]); side layout via fmov_side_offset(0..5) =
 * bias_sdf[8x256], b8[257], w8row[256], bias_col[4x256], bc4[3], wc4[3x256] */
int fmov_pack_all(const float* const* srcs, void* blob, float* side, int need_backward, void* stream);
int fmov_side_floats(void);
int fmov_side_offset(int which);
int fmov_fine_stash_count(void);
int fmov_fine_stash_blocks(int id);
int fmov_fine_stash_is_forward(int id);   /* 1: written by fmov_fine_fwd (a forward-only render needs only these) */
/* ge (fwd -> bwd) and eb_scratch (bwd) are opaque per-sample scratch buffers of 40 * 128 * ceil(B*S / 128) floats
 * (tile-major, [tile][40][128]) */
int fmov_fine_fwd(long long B, int S, const float* rays_o, const float* rays_d, const float* z, float sample_dist,
                  const void* wblob, void* const* stash, const float* bias_sdf, const float* b8, const float* w8row,
                  const float* bias_col, const float* bc4, float* sdf, float* nrm, float* rgb, float* ge, void* stream);
int fmov_fine_bwd(long long B, int S, const float* rays_o, const float* rays_d, const float* z, float sample_dist,
                  const void* wblob, void* const* stash, const float* bias_sdf, const float* b8, const float* w8row,
                  const float* bias_col, const float* bc4, const float* wc4, const float* rgb, const float* ge,
                  const float* d_sdf, const float* d_nrm, const float* d_rgb, const float* amax, float* d_pts, float* d_dirs,
                  float* zc4, float* eb_scratch, void* stream);
/* weight / bias gradients into one flat fp32 buffer (zeroed by the call); fmov_grad_offset(kind, layer):
 * kind 0 sdf weight [out,in], 1 sdf bias, 2 colour weight, 3 colour bias (effective weights, reference shapes) */
long long fmov_grad_offset(int kind, int layer);
long long fmov_grad_floats(void);
int fmov_dw(long long P, void* const* stash, const float* d_sdf, const float* zc4, const float* amax, float* grads,
            void* stream);
/* device scalar amax = max|upstream per-sample gradient|: sets the power-of-two loss scale of the fp16 gradient tiles */
int fmov_grad_amax(const float* d_sdf, const float* d_nrm, const float* d_rgb, long long P, float* amax, void* stream);
int fmov_grad_is_bf16(void);

/* ---- optimiser tail: gradient gather + Adam in two launches ------------------------------------------------
 * replaces optimizer.step() / pose_optimizers[k].step() (torch.optim.Adam; exp_runner.py:258-269, 812-816) and the
 * flatten / unflatten copies around the gradient all-reduce.  fmov_grad_gather: HOST arrays of n (<= 160) gradient tensors
 * -> flat G (+ one flag per parameter group at G[n_total..]); fmov_adam_step: Adam over every tensor of every group whose
 * flag is > 0, device tables built once by the caller (see csrc/optim.cu); lr / step [n_groups] are device floats.     */
int fmov_adam_chunk(void);
int fmov_grad_gather(int n, const float* const* src, const long long* dst_off, const int* numel, int n_groups,
                     unsigned long long active_mask, const float* flags_in, float* G, long long n_total, void* stream);
int fmov_adam_step(float* const* param, const long long* off, const int* numel, const int* group, const int* chunk_tensor,
                   const int* chunk_e0, int n_chunks, int n_groups, const float* G, long long n_total, float* M, float* V,
                   const float* lr, float* step, float beta1, float beta2, float eps, float grad_scale, unsigned int* done,
                   void* stream);

/* ---- marching cubes on the dense u grid -------------------------------------------------------------------
 * replaces `mcubes.marching_cubes(u, threshold)` (PyMCubes, CPU) at models/renderer.py:43 and the rescale of
 * models/renderer.py:47-50.  u: [X,Y,Z] fp32, z fastest (the layout fmov_sdf_query_grid writes).  Indexed mesh: one vertex
 * per crossed grid edge, ordered by (grid point x-major, axis); triangles ordered by (cell x-major, case-table order).
 *   1. fmov_mc_set_tables (once): HOST case table [256][15] + triangle counts [256] (fmov_pose_b200/mc_tables.py)
 *   2. fmov_mc_count: per 256-point chunk (fmov_mc_chunk_count of them) the number of vertices / triangles
 *   3. fmov_mc_scan: exclusive prefix sums (int64) over the chunks WITH the total appended (n_chunks + 1 entries: chunk c
 *      emits [off[c], off[c+1]), empty chunks are skipped); totals [2] = (V, T) size the outputs (the caller's one
 *      device->host read)
 *   4. fmov_mc_vertices: verts [V,3] = index coordinate * (sx,sy,sz) + (ox,oy,oz); vid3 [X*Y*Z,3] int32 scratch receives the
 *      vertex id of every crossed edge (other entries stay unwritten and are never read)
 *   5. fmov_mc_triangles: tris [T,3] int32 vertex ids                                                                    */
int fmov_mc_set_tables(const signed char* tri_table, const unsigned char* n_tris);
long long fmov_mc_chunk_count(int X, int Y, int Z);
long long fmov_mc_group_count(int X, int Y, int Z);          /* groups of 4096 chunks: entries of fmov_mc_scan's scratch */
/* list [n_chunks] int32 + n_list [1] int32 (device): fmov_mc_count appends the index of every chunk that emits something
 * (n_list is zeroed by the call, the order of the list is unspecified); the emit passes walk that list */
int fmov_mc_count(const float* u, int X, int Y, int Z, float iso, int* chunk_nv, int* chunk_nt, int* list, int* n_list,
                  void* stream);
/* group_scratch [fmov_mc_group_count] (device, no initialisation needed); chunk_voff / chunk_toff [n_chunks + 1]; totals [2] */
int fmov_mc_scan(const int* chunk_nv, const int* chunk_nt, long long n_chunks, unsigned long long* group_scratch,
                 long long* chunk_voff, long long* chunk_toff, long long* totals, void* stream);
int fmov_mc_vertices(const float* u, int X, int Y, int Z, float iso, const long long* chunk_voff, const int* list,
                     const int* n_list, float sx, float sy, float sz, float ox, float oy, float oz, float* verts, int* vid3,
                     void* stream);
int fmov_mc_triangles(const float* u, int X, int Y, int Z, float iso, const long long* chunk_toff, const int* list,
                      const int* n_list, const int* vid3, int* tris, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* FMOV_B200_H_ */
