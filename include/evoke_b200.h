/* evoke_b200.h — C ABI of libevoke_b200.so: B200 (sm_100a) kernels for EVOKE's multi-view,
 * multi-positive image-text contrastive objective, forward and backward.
 *
 * What this replaces.  The reference has no FFI for this path: the objective is two Python
 * methods of the `Pretrain` nn.Module that run as stock PyTorch eager ops
 * (models/model_pretrain_finetune_v0520.py, same bodies in the five sibling model files):
 *     global_alignment_loss(self, global_image_embed, global_text_embed, patient_ids)  :486-504
 *     multi_pos_contra_images_v0401(self, global_image_embed, patient_ids)             :421-446
 * The drop-in (python package `evoke_b200`) keeps those signatures and
 * lowers them onto the entry points below, loaded with ctypes.  Each entry point names the
 * reference lines it stands in for.
 *
 * Conventions
 *  - Plain pointers and sizes only.  Every pointer is a DEVICE pointer unless it says host.
 *  - The library never allocates, frees or retains device memory, never synchronises the
 *    device, and enqueues on the caller's `stream` (a cudaStream_t passed as void*): calls
 *    are CUDA-graph capturable and re-entrant.  Workspaces are sized by the evk_*_bytes helpers
 *    below and owned by the caller.  (One exception, stated there: evk_peer_alloc, because
 *    CUDA-IPC handles name whole cudaMalloc allocations.)
 *  - Return value: EVK_OK or a negative code; evk_last_error() returns a thread-local,
 *    human-readable message for the last failing call on this thread.  Launch errors are
 *    picked up with cudaGetLastError(), without a device sync.
 *  - There is no CPU fallback.  On a device that is not sm_100 the tcgen05 entry points
 *    return EVK_ERR_UNSUPPORTED.
 *
 * Notation:  M_ij = [id_i == id_j],  c_i = sum_j M_ij,  Qhat/Khat = L2-normalised rows,
 * S = Qhat Khat^T * inv_tau,  shift = inv_tau (|S| <= inv_tau for unit rows),
 * E_ij = exp(S_ij - shift).
 */
#ifndef EVOKE_B200_H_
#define EVOKE_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EVK_OK                0
#define EVK_ERR_INVALID      -1   /* bad argument (shape, alignment, null pointer) */
#define EVK_ERR_CUDA         -2   /* CUDA runtime / driver error, see evk_last_error() */
#define EVK_ERR_UNSUPPORTED  -3   /* device or configuration not supported (no fallback) */

#define EVK_DTYPE_F32   0
#define EVK_DTYPE_BF16  1
#define EVK_DTYPE_F16   2

/* flags for the mpce entry points */
#define EVK_FLAG_EXCLUDE_DIAG  1  /* column (row + diag_offset) is removed from the softmax:
                                     multi_pos_contra_images_v0401 :438 (fill_diagonal_(-1e9)) */
#define EVK_FLAG_NO_COLSUM     2  /* skip column sums (symmetric problem: MPC) */
#define EVK_FLAG_SPLIT_BF16    4  /* operands are (hi, lo) bf16 pairs: S = hi.hi + hi.lo + lo.hi,
                                     ~2^-17 relative, the fp32-parity mode */
#define EVK_FLAG_NO_POS        8  /* K3 leaves the positive-logit sums to evk_mpce_pos (bits may be NULL) */
#define EVK_FLAG_AVGPOS       16  /* 'averaged positive logit' rule of PretrainNewMulPos (:748-815, :670-708): the
                                     positives of a row enter the softmax as ONE logit, their mean (small path) */
#define EVK_FLAG_PUBLIC_MASK  31  /* every entry point ignores bits outside the EVK_FLAG_* set */
/* Every loss entry point requires 0 < inv_tau <= EVK_MAX_INV_TAU (tau >= 0.025): the softmax uses the fixed
 * shift 1/tau (unit rows: |S| <= 1/tau) and exp(-2/tau) must stay a normal fp32 number. */
#define EVK_MAX_INV_TAU     40.0f

typedef void* evk_stream_t;

/* Cross-GPU synchronisation folded into the head of a consumer kernel of the sharded path (HOST struct, read
 * during the call).  The kernel publishes what this GPU stored into peer memory in EARLIER kernels of the stream
 * (system fence + release store of the epoch into every rank's flag area) and waits until every peer has done the
 * same, i.e. it is evk_peer_barrier without the extra launch.  epoch = per_step * (*step) + index: `step` is the
 * transport's device step counter (advanced by evk_shard_prologue), `index` in 1..per_step numbers the sync points
 * of one step.  flag_ptrs[t] / err_ptrs[t]: rank t's flag area (>= 16 uint32, zeroed at start, used by these
 * folded syncs only) and failure flag, peer-mapped.  Timeout semantics as evk_peer_barrier. */
typedef struct evk_peer_sync {
  uint64_t flag_ptrs[16];
  uint64_t err_ptrs[16];
  int* err_host;
  const int* step;
  int n_ranks, rank, index, per_step;
  int64_t timeout_ms;
} evk_peer_sync_t;

#if defined(__GNUC__)
#define EVK_API __attribute__((visibility("default")))
#else
#define EVK_API
#endif

EVK_API int         evk_version(void);                 /* ABI version, bumped on any signature change */
EVK_API const char* evk_last_error(void);
/* host query; sm_count / cc may be NULL */
EVK_API int         evk_device_info(int device, int* sm_count, int* cc_major, int* cc_minor);

/* ---- sizes of caller-owned buffers (host functions; <= 0 means bad arguments) ------------------------------
 * workspace of evk_mpce_stats_fused / evk_mpce_shard_stats_push for n_rows rows and n_cols columns (0 if no
 * column statistics) */
EVK_API int64_t evk_stats_workspace_bytes(int64_t n_rows, int64_t n_cols);
/* workspace of evk_mpce_shard_finish */
EVK_API int64_t evk_shard_finish_workspace_bytes(int64_t n_cols);
/* row pitch (uint32 words) of the K2 mask the tcgen05 entry points expect: whole 256-column tiles, 16-byte rows */
EVK_API int64_t evk_posmask_ld_words(int64_t n_cols);
/* rows of the row-statistic partial buffers (rs_part / rp_part) K3 writes for n_cols key columns */
EVK_API int64_t evk_mpce_rowpart_rows(int64_t n_cols);
/* rows of the column-statistic partial buffer (cs_part) K3 writes for n_rows query rows */
EVK_API int64_t evk_mpce_colpart_rows(int64_t n_rows);
/* row pitch (bf16 elements) of the E / W strip for n_cols key columns */
EVK_API int64_t evk_mpce_strip_ld(int64_t n_cols);

/* ---- K1: fused L2-normalise (+ bf16 hi/lo split, + row gather) --------------------------
 * Replaces F.normalize(x, dim=-1, p=2) at :495-496 and :436: xhat = x / max(||x||_2, 1e-12).
 * x is [*, d] with arbitrary element strides (the reference passes the strided slice
 * `[:,0,:]` of the permuted projection-head output, :484/:399).  Output row r is computed from
 * input row gather[r] (gather == NULL: r), which implements the row filter of :426-429.
 * Any of out_f32 / out_hi / out_lo may be NULL.  out_hi = bf16(xhat), out_lo = bf16(xhat - hi).
 * ld_* are row pitches in elements; ld_bf16 must be a multiple of 8 (16-byte rows for TMA).
 * norm[r] = ||x_row||_2 (unclamped; the backward needs to know whether the clamp was active). */
EVK_API int evk_l2norm_fwd(const void* x, int x_dtype, int64_t n_out, int64_t d,
                   int64_t stride_row, int64_t stride_col, const int32_t* gather,
                   float* out_f32, int64_t ld_f32,
                   void* out_hi, void* out_lo, int64_t ld_bf16,
                   float* norm, evk_stream_t stream);

/* Backward of K1 fused with the upstream-gradient scale:
 *   dx_row = scale * (g - xhat (xhat.g)) / ||x||      (||x|| >= 1e-12)
 *          = scale * g / 1e-12                         (clamp active)
 * scale = scale_host * (scale_dev ? *scale_dev : 1).  g is fp32 [n_out, ld_g] (dXhat).
 * dx is written at row gather[r] (or r) with pitch ld_dx, dtype dx_dtype; rows that are not
 * written are the caller's to zero.  If accumulate != 0, dx += (fp32 only). */
EVK_API int evk_l2norm_bwd(const void* x, int x_dtype, int64_t n_out, int64_t d,
                   int64_t stride_row, int64_t stride_col, const int32_t* gather,
                   const float* norm, const float* g, int64_t ld_g,
                   const float* scale_dev, float scale_host,
                   void* dx, int dx_dtype, int64_t ld_dx, int accumulate, evk_stream_t stream);

/* Same, with the upstream gradient given as n_parts partial buffers that are summed on the fly (in index
 * order): g_total[r, c] = sum_p g[p * part_stride + r * ld_g + c] (g_dtype: EVK_DTYPE_F32 or _BF16).  Closes the fused reduce-scatter of
 * evk_mpce_bwd_gemm_scatter(store = 1): part p is what rank p's contraction stored for this rank's rows.
 * error (may be NULL): device int; when it is non-zero (a cross-GPU barrier of this step timed out,
 * evk_peer_barrier) every gradient written is NaN instead of a silently wrong value.
 * sync (may be NULL): see evk_peer_sync_t - the partial buffers were stored by the peers' contractions. */
EVK_API int evk_l2norm_bwd_parts(const void* x, int x_dtype, int64_t n_out, int64_t d,
                         int64_t stride_row, int64_t stride_col, const int32_t* gather,
                         const float* norm, const void* g, int g_dtype, int64_t ld_g, int n_parts, int64_t part_stride,
                         const float* scale_dev, float scale_host,
                         void* dx, int dx_dtype, int64_t ld_dx, int accumulate, const int* error,
                         const evk_peer_sync_t* sync, evk_stream_t stream);

/* ---- K2: positive-mask builder -------------------------------------------------------------
 * Replaces (ids.reshape(-1,1) == ids.reshape(1,-1)) + .float().to(device) + rowsum at
 * :488-491 and :422-424/:430.  bits[r, w] bit k = [key(row r) == key(col 32w+k)], i.e. exactly
 * np.packbits(M, axis=1, bitorder='little') as little-endian uint32; counts[r] = popcount of
 * row r = c_r.  A key is id, or the pair (id, id2) when the id2 pointers are non-NULL
 * (patient AND study, the reference's "p<subject>_s<study>" string, dataloaders_v0401.py:83).
 * With clear_diag the bit at column (r + diag_offset) is cleared (:424).  All ld_words words
 * of every row are written (bits at columns >= n_cols are zero); ld_words >= ceil(n_cols/32).
 * bits == NULL (then pos_idx is required): counts and lists only - nothing of size N^2 is written; the bf16
 * mode of the large path runs on those (evk_mpce_pos_from_lists, evk_mpce_w_from_e).
 * pos_idx (may be NULL): [n_rows, pos_slots] int32, pos_idx[r, s] = column of the s-th positive of row r for
 * s < min(counts[r], pos_slots), unspecified order, other entries untouched: the sparse form of the same mask
 * for the O(N) consumers (evk_mpce_pos_logits, evk_mpce_w_from_e).
 * counts_zeroed != 0: the caller guarantees counts is all zero on entry (evk_shard_prologue zeroes it), so no
 * memset node is enqueued.  sync (may be NULL): see evk_peer_sync_t - the column ids were stored by the peers. */
EVK_API int evk_posmask_build(const int32_t* ids_row, const int32_t* ids2_row, int64_t n_rows,
                      const int32_t* ids_col, const int32_t* ids2_col, int64_t n_cols,
                      int64_t diag_offset, int clear_diag,
                      uint32_t* bits, int64_t ld_words, int32_t* counts,
                      int32_t* pos_idx, int pos_slots, int counts_zeroed, const evk_peer_sync_t* sync,
                      evk_stream_t stream);

/* ---- small path: fp32 SIMT fused kernels (reference-sized batches, N <~ 1k) ---------------
 * One launch handles the rows of `q` against all columns `k` (both already normalised, fp32):
 * forward:  row_sum[i] = sum_j E_ij (diagonal excluded if flagged),
 *           row_pos[i] = sum_j M_ij S_ij
 * replacing the mm + `/temp` + log_softmax of :499-502 / :437-443 for one direction; the
 * caller runs it once per direction.  Nothing of size n_rows x n_cols is ever stored.
 * With EVK_FLAG_AVGPOS row_sum[i] runs over the NEGATIVES only. */
EVK_API int evk_mpce_small_fwd(const float* q, int64_t ld_q, const float* k, int64_t ld_k,
                       int64_t n_rows, int64_t n_cols, int64_t d,
                       const uint32_t* bits, int64_t ld_words,
                       float inv_tau, int flags, int64_t diag_offset,
                       float* row_sum, float* row_pos, evk_stream_t stream);

/* backward for the same rows:  dq[i,:] = sum_j W_ij k[j,:],
 *   W_ij = E_ij (a_i + b_j) - 2 M_ij / c_i      (0 on an excluded diagonal)
 * i.e. N*tau*(dS + its transpose-direction term) of the closed form in oracle/evoke_oracle.py;
 * the 1/(2N tau) (or 1/(M' tau)) factor is applied by evk_l2norm_bwd's scale.
 * With EVK_FLAG_AVGPOS (then counts may be NULL and pos_row / pos_col are required, else they are ignored):
 *   W_ij = E_ij (a_i + b_j) for negatives,  pos_row[i] + pos_col[j] for positives,
 * a / pos_* from evk_mpce_finalize_avgpos of the row resp. column direction. */
EVK_API int evk_mpce_small_bwd(const float* q, int64_t ld_q, const float* k, int64_t ld_k,
                       int64_t n_rows, int64_t n_cols, int64_t d,
                       const uint32_t* bits, int64_t ld_words, const int32_t* counts,
                       const float* a_row, const float* b_col,
                       float inv_tau, int flags, int64_t diag_offset,
                       float* dq, int64_t ld_dq,
                       const float* pos_row, const float* pos_col, evk_stream_t stream);

/* Batched forms (independent problems of one shape, e.g. one per sample): problem b uses q + b*bs_q, k + b*bs_k
 * and the row / column vectors (row_sum, row_pos, a_row, b_col) at offset b*bs_vec, dq at b*bs_dq; the mask and
 * counts are shared by all problems.  Used for the per-sample token-level InfoNCE of
 * Pretrain.local_text_token_alignment_loss (:518-525), which is the G loss with identity ids on every sample. */
EVK_API int evk_mpce_small_fwd_batched(const float* q, int64_t ld_q, int64_t bs_q, const float* k, int64_t ld_k, int64_t bs_k,
                               int64_t batch, int64_t n_rows, int64_t n_cols, int64_t d,
                               const uint32_t* bits, int64_t ld_words, float inv_tau, int flags,
                               float* row_sum, float* row_pos, int64_t bs_vec, evk_stream_t stream);
EVK_API int evk_mpce_small_bwd_batched(const float* q, int64_t ld_q, int64_t bs_q, const float* k, int64_t ld_k, int64_t bs_k,
                               int64_t batch, int64_t n_rows, int64_t n_cols, int64_t d,
                               const uint32_t* bits, int64_t ld_words, const int32_t* counts,
                               const float* a_row, const float* b_col, int64_t bs_vec,
                               float inv_tau, int flags, float* dq, int64_t ld_dq, int64_t bs_dq, evk_stream_t stream);

/* ---- f1: Pretrain.local_text_token_alignment_loss (:506-526), the parameter-free cross-attention ------------
 * text [batch, l, d], image [batch, p, d], contiguous fp32.  Forward (:509-511):
 *   att[b, i, :] = softmax_p( text[b, i] . image[b, p] / sqrt(d) ),   out[b, i] = sum_p att[b, i, p] image[b, p]
 * Backward, given d_out = dL/d out: d_image is written, d_text is ACCUMULATED into (it already holds the gradient
 * that reaches the text tokens through their own normalisation, :515); ds is a [batch, l, p] workspace.
 * fp32 SIMT, one CTA per (8 tokens, sample); deterministic.  p <= 1024, d <= 4096, 8 (d + p) floats of shared memory. */
EVK_API int evk_local_attend_fwd(const float* text, const float* image, int64_t batch, int64_t l, int64_t p, int64_t d,
                         float* att, float* out, evk_stream_t stream);
EVK_API int evk_local_attend_bwd(const float* text, const float* image, const float* att, const float* d_out,
                         int64_t batch, int64_t l, int64_t p, int64_t d,
                         float* ds, float* d_text, float* d_image, evk_stream_t stream);

/* Per-sample token-level InfoNCE of f1 (:518-525) for l <= 128 tokens, register-blocked fp32, one sample per CTA:
 *   e_out[b, i, j] = exp(th[b,i] . oh[b,j] * inv_tau - inv_tau),  row_sum / col_sum = its row / column sums per sample
 *   (vectors of batch*l), row_pos[b*l + i] = the diagonal logit (identity targets :520).
 * th / oh: the L2-normalised text / attended tokens, contiguous [batch*l, d].  evk_mpce_finalize turns the sums into
 * a_row / b_col / loss; the backward then is d_th = W oh, d_oh = W^T th with W = E (a_i + b_j) - 2 [i == j]. */
EVK_API int evk_token_sim_fwd(const float* th, const float* oh, int64_t batch, int64_t l, int64_t d, float inv_tau,
                      float* e_out, float* row_sum, float* row_pos, float* col_sum, evk_stream_t stream);
EVK_API int evk_token_sim_bwd(const float* th, const float* oh, const float* e_in, const float* a_row, const float* b_col,
                      int64_t batch, int64_t l, int64_t d, float* d_th, float* d_oh, evk_stream_t stream);

/* ---- statistics -> loss ----------------------------------------------------------------------
 * out[j] = sum_p part[p*ld + j], p < parts: deterministic reduction of per-tile partials;
 * if divisor != NULL the sum is divided by divisor[j] (0 where divisor[j] <= 0): pos_j / c_j. */
EVK_API int evk_reduce_partials(const float* part, int64_t parts, int64_t ld, int64_t n,
                        const int32_t* divisor, float* out, evk_stream_t stream);

/* Turns the O(N) statistics into the loss and the backward's scale vectors.
 *   a_row[i] = 1/row_sum[i];  b_col[j] = 1/col_sum[j]  (col_sum may be NULL: then b_col = NULL ok)
 *   loss_out[0] = inv_count * [ sum_i (shift + ln row_sum[i] - pos_weight * row_pos[i]/c_i)
 *                             + sum_{j in [col_lo,col_hi)} (shift + ln col_sum[j]) ]
 * G loss (:501-503): inv_count = 1/(2N), pos_weight = 2 (the two directions share their
 * positive term because M_ij=1 implies c_i=c_j).  MPC (:443): inv_count = 1/M', pos_weight = 1,
 * no column part.  loss_out is a single fp32; the sum is carried in fp64. */
EVK_API int evk_mpce_finalize(const float* row_sum, const float* row_pos, const int32_t* counts,
                      int64_t n_rows, const float* col_sum, int64_t n_cols,
                      int64_t col_lo, int64_t col_hi, float shift, float pos_weight,
                      double inv_count, float* a_row, float* b_col, float* loss_out,
                      evk_stream_t stream);

/* One softmax direction of the 'averaged positive logit' rule (PretrainNewMulPos :783-811, v0404 :691-705),
 * from evk_mpce_small_fwd(EVK_FLAG_AVGPOS)'s row_neg / row_pos:
 *   pbar_i = row_pos[i]/c_i, u = exp(pbar_i - shift), Z = u + row_neg[i]
 *   a_row[i] = 1/Z, pos_row[i] = (u/Z - 1)/c_i, loss_out[0] (+)= inv_count * sum_i (shift - pbar_i + ln Z)
 * rows with c_i = 0 contribute nothing (a = pos = 0).  accumulate != 0 adds to loss_out (second direction). */
EVK_API int evk_mpce_finalize_avgpos(const float* row_neg, const float* row_pos, const int32_t* counts,
                             int64_t n_rows, float shift, double inv_count,
                             float* a_row, float* pos_row, float* loss_out, int accumulate,
                             evk_stream_t stream);

/* Single-GPU fused form of evk_reduce_partials (x3) + evk_mpce_finalize: takes the per-tile
 * partials of K3 directly (rs_part: [row_parts, ld_row]; rp_part: [pos_parts, ld_pos] - K3's
 * partials, or the single row written by evk_mpce_pos; cs_part: [col_parts, ld_col] or NULL), the loss covers all rows and all columns.  workspace: >= 16 + 24*ceil(max(n_rows,n_cols)/32)
 * bytes, 16-byte aligned, contents irrelevant.  Deterministic (fixed summation order). */
EVK_API int evk_mpce_stats_fused(const float* rs_part, int64_t row_parts, int64_t ld_row,
                         const float* rp_part, int64_t pos_parts, int64_t ld_pos,
                         const int32_t* counts, int64_t n_rows,
                         const float* cs_part, int64_t col_parts, int64_t ld_col, int64_t n_cols,
                         int64_t col_lo, int64_t col_hi,
                         float shift, float pos_weight, double inv_count,
                         float* a_row, float* b_col, float* loss_out,
                         void* workspace, int64_t workspace_bytes, evk_stream_t stream);

/* ---- large path: tcgen05 / TMEM / TMA kernels (sm_100a only) ------------------------------
 * All operands are bf16 row-major with 16-byte aligned base and row pitch (ld % 8 == 0).
 * With EVK_FLAG_SPLIT_BF16 the *_lo pointers carry the low halves written by K1.
 *
 * K3 forward (flash-style: S tiles live only in TMEM):
 *   row_sum_part[P*cb + part, i] / row_pos_part[...] : partial over part `part` of column block cb (256 columns),
 *                                               P = evk_mpce_row_parts() (one per epilogue warp sharing a row):
 *                                               P * ceil(n_cols/256) partial rows
 *   col_sum_part[rb, j]                       : partial over row block rb (128 rows)
 * pitches: ld_rowpart >= n_rows, ld_colpart >= n_cols.  Reduce with evk_reduce_partials. */
EVK_API int evk_mpce_row_parts(void);
EVK_API int evk_mpce_fwd(const void* q_hi, const void* q_lo, int64_t ld_q,
                 const void* k_hi, const void* k_lo, int64_t ld_k,
                 int64_t n_rows, int64_t n_cols, int64_t d,
                 const uint32_t* bits, int64_t ld_words,
                 float inv_tau, int flags, int64_t diag_offset,
                 float* row_sum_part, float* row_pos_part, int64_t ld_rowpart,
                 float* col_sum_part, int64_t ld_colpart, evk_stream_t stream);

/* K3 that ALSO stores E_ij = exp(S_ij - inv_tau) as bf16 in the row strip e_out [n_rows, ld_e]
 * (ld_e % 8 == 0, ld_e >= n_cols; entries outside the softmax - the excluded diagonal - are 0).
 * bf16 mode only (no split operands).  With the strip the backward needs no second sweep over the
 * similarity tiles: evk_mpce_w_from_e turns E into W in place and the step executes 6 N^2 D FLOP
 * instead of 8 N^2 D.  Same statistics outputs as evk_mpce_fwd. */
EVK_API int evk_mpce_fwd_store(const void* q_hi, int64_t ld_q, const void* k_hi, int64_t ld_k,
                       int64_t n_rows, int64_t n_cols, int64_t d,
                       const uint32_t* bits, int64_t ld_words,
                       float inv_tau, int flags, int64_t diag_offset,
                       float* row_sum_part, float* row_pos_part, int64_t ld_rowpart,
                       float* col_sum_part, int64_t ld_colpart,
                       void* e_out, int64_t ld_e, evk_stream_t stream);

/* Raw logits of the listed positives: pos_dot[i, s] = q_i . k_{pos_idx[i, s]} (bf16 operands, fp32 accumulate)
 * for s < min(counts[i], pos_slots).  O(N*D); meant for a side stream next to K3.  evk_mpce_w_from_e turns
 * them into exact W entries without touching the mask or the operands again. */
EVK_API int evk_mpce_pos_logits(const void* q_hi, int64_t ld_q, const void* k_hi, int64_t ld_k,
                        int64_t n_rows, int64_t d, const int32_t* pos_idx, const int32_t* counts, int pos_slots,
                        float* pos_dot, evk_stream_t stream);

/* Positive-logit sums WITHOUT the dense mask, for use with EVK_FLAG_NO_POS in bf16 mode:
 *   row_pos[i] = inv_tau * sum_{j in P_i} q_i . k_j   (the sum_j Y_ij S_ij term of :501-502 / :443 before the 1/c_i)
 * from the K2 lists (pos_dot of evk_mpce_pos_logits) when counts[i] <= pos_slots, else by scanning the column
 * ids for the row's key (clear_diag / diag_offset as in K2) and computing the dot products.  O(N) + O(D) per
 * positive; nothing of size N^2 is read. */
EVK_API int evk_mpce_pos_from_lists(const void* q_hi, int64_t ld_q, const void* k_hi, int64_t ld_k,
                            int64_t n_rows, int64_t n_cols, int64_t d,
                            const int32_t* ids_row, const int32_t* ids2_row,
                            const int32_t* ids_col, const int32_t* ids2_col, int64_t diag_offset, int clear_diag,
                            const int32_t* counts, const float* pos_dot, int pos_slots, float inv_tau,
                            float* row_pos, evk_stream_t stream);

/* Sharded K3 that starts before the all-gather of the key rows has finished.  k_hi is this rank's buffer of ALL
 * key rows, filled by every rank's evk_peer_push_shard while the sweep runs: column c belongs to source
 * c / cols_per_source, and landed[s] (this rank's landed-counter area) reaches landed_per_step * *step once
 * source s's rows are complete (landed_per_step = the n_ctas of evk_peer_push_shard).  The column blocks are visited starting at first_col (this rank's own rows, already in place) and the
 * TMA producer waits for a source's flag before its first load from that source's columns; if a flag does not
 * arrive within 2 s *error is set and the sweep continues (a dead peer must not hang the GPU).
 * e_out may be NULL (statistics only).  Otherwise as evk_mpce_fwd_store. */
EVK_API int evk_mpce_fwd_store_gathered(const void* q_hi, int64_t ld_q, const void* k_hi, int64_t ld_k,
                                int64_t n_rows, int64_t n_cols, int64_t d,
                                const uint32_t* bits, int64_t ld_words,
                                float inv_tau, int flags, int64_t diag_offset,
                                float* row_sum_part, float* row_pos_part, int64_t ld_rowpart,
                                float* col_sum_part, int64_t ld_colpart,
                                void* e_out, int64_t ld_e,
                                const uint32_t* landed, const int* step, int* error,
                                int64_t cols_per_source, int64_t first_col, int landed_per_step, evk_stream_t stream);

/* K4t, in place over the strip written by evk_mpce_fwd_store (or any row range of it: offset the
 * strip / bits / counts / a_row pointers):  strip[i, j] <- bf16( E_ij (a_row[i] + b_col[j]) - 2 M_ij / c_i ),
 * the W of evk_mpce_small_bwd; HBM-bound (4 bytes per (i, j) + 1 mask bit).  MPC: pass b_col = a_row.
 * q_hi / k_hi (the bf16 operands of the forward, K1's padded rows; d, inv_tau as in the forward) let the
 * kernel recompute the POSITIVE entries from S_ij in fp32 instead of from the bf16-rounded E: that is
 * where softmax and target cancel, and it keeps cold temperatures inside the bf16-mode tolerance.
 * q_hi == NULL skips this (all entries from the strip).  With pos_idx / pos_dot (K2's lists and
 * evk_mpce_pos_logits' values, [n_rows, pos_slots]) rows with at most pos_slots positives take their exact
 * entries from the lists - no mask scan, no dot products in the backward; the other rows scan the mask.
 * bits == NULL (no dense mask was built: evk_posmask_build with bits == NULL): the lists, the operands and the
 * ids are required, and rows with more than pos_slots positives find them by scanning the column ids
 * (ids_row [n_rows] / ids_col [n_cols], optional second key component, clear_diag / diag_offset as in K2). */
EVK_API int evk_mpce_w_from_e(void* strip, int64_t ld_e, int64_t n_rows, int64_t n_cols,
                      const uint32_t* bits, int64_t ld_words, const int32_t* counts,
                      const float* a_row, const float* b_col,
                      const void* q_hi, int64_t ld_q, const void* k_hi, int64_t ld_k, int64_t d, float inv_tau,
                      const int32_t* pos_idx, const float* pos_dot, int pos_slots,
                      const int32_t* ids_row, const int32_t* ids2_row, const int32_t* ids_col, const int32_t* ids2_col,
                      int64_t diag_offset, int clear_diag, evk_stream_t stream);

/* Positive-logit sums from the bit mask, for use with EVK_FLAG_NO_POS:
 *   row_pos[i] = sum_j M_ij S_ij = inv_tau * sum_{j: bit (i,j) set} q_i . k_j
 * (the sum_j Y_ij S_ij term of the soft-target CE, :501-502 / :443, before the 1/c_i).  O(N*D) work;
 * meant to run on a side stream concurrently with evk_mpce_fwd.  q_lo/k_lo: both NULL or both set. */
EVK_API int evk_mpce_pos(const void* q_hi, const void* q_lo, int64_t ld_q,
                 const void* k_hi, const void* k_lo, int64_t ld_k,
                 int64_t n_rows, int64_t n_cols, int64_t d,
                 const uint32_t* bits, int64_t ld_words, float inv_tau,
                 float* row_pos, evk_stream_t stream);

/* K4a backward, pass 1: recompute S tiles, form W (see evk_mpce_small_bwd) and store it as
 * bf16 (w_hi, and w_lo = bf16(W - hi) with EVK_FLAG_SPLIT_BF16) in a row strip
 * [n_rows, ld_w], ld_w % 8 == 0 and ld_w >= n_cols. */
EVK_API int evk_mpce_bwd_w(const void* q_hi, const void* q_lo, int64_t ld_q,
                   const void* k_hi, const void* k_lo, int64_t ld_k,
                   int64_t n_rows, int64_t n_cols, int64_t d,
                   const uint32_t* bits, int64_t ld_words, const int32_t* counts,
                   const float* a_row, const float* b_col,
                   float inv_tau, int flags, int64_t diag_offset,
                   void* w_hi, void* w_lo, int64_t ld_w, evk_stream_t stream);

/* K4b backward, pass 2: the two gradient contractions over the W strip, fp32 accumulate in
 * TMEM, fp32 atomic accumulation into `out` (split-K: the caller zeroes `out` first):
 *   transpose_w == 0:  out[i, :] += alpha * sum_j W[i, j] x[j, :]    (dQhat; out is [n_rows, d])
 *   transpose_w == 1:  out[j, :] += alpha * sum_i W[i, j] x[i, :]    (dKhat; out is [n_cols, d])
 * x is the bf16 normalised matrix of the OTHER side (hi, lo). */
EVK_API int evk_mpce_bwd_gemm(const void* w_hi, const void* w_lo, int64_t ld_w,
                      int64_t n_rows, int64_t n_cols, int transpose_w,
                      const void* x_hi, const void* x_lo, int64_t ld_x, int64_t d,
                      float alpha, int flags, float* out, int64_t ld_out, int cta_limit, evk_stream_t stream);

/* ---- peer-memory transport for the sharded path (NVLink / NVSwitch) --------------------------
 * *_ptrs are HOST arrays of device base addresses: the per-rank buffers of one symmetric
 * allocation (peer-mapped pointers, the local rank included).  Ordering across ranks: evk_peer_sync_t folded into
 * the consumer kernels, or evk_peer_barrier between producer and consumer kernels. */

/* Prologue of a sharded step in one launch: K1 of this rank's key rows (text) stored as bf16 at rows
 * row_offset.. of EVERY rank's key buffer (khat_ptrs: host table, n_dst entries), K1 of its query rows (image)
 * into the local q_hi, the id shard(s) pushed to offset row_offset of every rank's id buffer(s), and - if
 * zero_buf != NULL - the zero fill of an fp32 [n_rows, ld_zero] accumulator (the split-K output of the local
 * gradient contraction).  text / image: fp32, bf16 or fp16 with arbitrary element strides (the reference's
 * [:,0,:] head views, :484/:399; contiguous 16-byte aligned fp32 rows take 128-bit loads); d % 8 == 0,
 * d <= 2048.  F.normalize of :495-496 for both
 * sides plus what a sharded run must exchange before the similarity sweep.  n_dst = 1 with only this rank's
 * own buffer keeps the key rows local (evk_peer_push_shard then moves them next to the sweep); the ids always go
 * to all n_ids_dst ranks.  step_counter (may be NULL): device int advanced by one per launch - the epoch of the
 * landed flags.  zero_i32 (may be NULL): n_zero_i32 int32 to zero (K2's counts).  error (may be NULL): the transport's sticky failure flag (evk_peer_barrier); once set the
 * kernel writes nothing, in particular nothing into peer memory. */
EVK_API int evk_shard_prologue(const void* text, int text_dtype, int64_t text_stride, int64_t text_col_stride,
                       const void* image, int image_dtype, int64_t image_stride, int64_t image_col_stride,
                       int64_t n_rows, int64_t d, int n_dst, const uint64_t* khat_ptrs, int64_t ld_bf16,
                       int64_t row_offset, float* k_norm, void* q_hi, float* q_norm,
                       const int32_t* ids, const int32_t* ids2, int n_ids_dst,
                       const uint64_t* ids_ptrs, const uint64_t* ids2_ptrs,
                       float* zero_buf, int64_t ld_zero, int32_t* zero_i32, int64_t n_zero_i32,
                       int* step_counter, const int* error, evk_stream_t stream);

/* The all-gather of the key rows, overlapped with the similarity sweep.  Copies this rank's shard (`bytes` at
 * `src`, normally its own rows inside its own buffer) to byte offset dst_offset_bytes of every OTHER rank's
 * buffer, one destination at a time in the order rank+1, rank+2, ... (every GPU then receives from one source at
 * a time, at full NVLink rate, and the shards land in a known order).  n_ctas CTAs of ONE driving thread each move
 * 8 KiB chunks with TMA bulk copies (global -> shared -> peer): the copy takes no issue slots or registers from the
 * sweep that runs beside it.  A CTA that is done with a destination adds 1 (system scope, after its stores have
 * completed) to that destination's landed counter of this source, landed_ptrs[t][rank]; the counters are monotonic
 * over the steps (zeroed once at allocation): source s has landed in step k when landed[s] >= n_ctas * k.  Its own
 * counter is advanced by n_ctas at once. */
EVK_API int evk_peer_push_shard(const void* src, int64_t bytes, int n_ranks, int rank,
                        const uint64_t* dst_ptrs, int64_t dst_offset_bytes,
                        const uint64_t* landed_ptrs, int n_ctas, evk_stream_t stream);

/* Waits (one tiny kernel) until landed[s] >= per_step * *step for every source s < n_ranks: for consumers of the
 * gathered rows other than K3.  Sets *error after timeout_ms (<= 0: 2000) instead of hanging. */
EVK_API int evk_peer_wait_landed(const void* landed, int n_ranks, const int* step, int per_step, int* error,
                         int64_t timeout_ms, evk_stream_t stream);

/* Sharded form of evk_mpce_stats_fused: reduces K3's partials of this rank's row block, writes a_row, and
 * stores this rank's statistics slot - the raw partial column sums (n_cols floats) followed by its row-side
 * loss term inv_count * sum_i (shift + ln R_i - pos_weight pos_i / c_i) - at element offset slot_offset of
 * EVERY rank's slot buffer (slot_ptrs: host table of n_dst peer-mapped addresses).  evk_mpce_shard_finish
 * closes the forward after a barrier.  workspace as for evk_mpce_stats_fused; workspace_persistent as for
 * evk_mpce_shard_finish. */
EVK_API int evk_mpce_shard_stats_push(const float* rs_part, int64_t row_parts, int64_t ld_row,
                              const float* rp_part, int64_t pos_parts, int64_t ld_pos,
                              const int32_t* counts, int64_t n_rows,
                              const float* cs_part, int64_t col_parts, int64_t ld_col, int64_t n_cols,
                              float shift, float pos_weight, double inv_count, float* a_row,
                              const uint64_t* slot_ptrs, int n_dst, int64_t slot_offset,
                              void* workspace, int64_t workspace_bytes, int workspace_persistent, evk_stream_t stream);

/* Symmetric buffers for that transport.  evk_peer_alloc is the ONE place the library allocates device
 * memory (cudaMalloc + cudaMemset, i.e. it also synchronises; called once per transport context, never per step): CUDA-IPC handles name whole allocations, so the exchanged buffers
 * cannot come out of a caching allocator.  The caller frees them with evk_peer_free.  evk_peer_export
 * writes the 64-byte IPC handle of such a buffer to HOST memory; evk_peer_open maps a peer's handle
 * (received through any host channel, e.g. torch.distributed) into this process; evk_peer_close unmaps. */
EVK_API int evk_peer_alloc(int64_t bytes, void** ptr_out);
EVK_API int evk_peer_free(void* ptr);
EVK_API int evk_peer_export(const void* ptr, void* handle_out_64_bytes);
EVK_API int evk_peer_open(const void* handle_64_bytes, void** ptr_out);
EVK_API int evk_peer_close(void* ptr);

/* Barrier across the ranks' GPUs, enqueued on `stream` (one tiny kernel, CUDA-graph capturable).
 * flag_ptrs: HOST array of n_ranks device addresses, entry t = rank t's flag area (>= 16 uint32, zeroed at
 * start, peer-mapped).  epoch: this rank's device-resident uint32 counter (zeroed at start; advanced by the
 * kernel).  Everything this rank wrote to peer memory in earlier kernels of the stream is visible to the peers
 * once they pass.  error_ptrs: HOST array of n_ranks device addresses, entry t = rank t's failure flag (an int in
 * the same symmetric, peer-mapped memory; zeroed at start).  If a peer does not arrive within timeout_ms
 * (<= 0: 2000) the failure flag of EVERY rank is set to 1 and the kernel returns: a dead peer must not hang the
 * GPU, and the late peer - which will pass its own barriers at once - must not trust this rank's buffers either.
 * The flag is sticky and is what makes the failure loud without a host sync: evk_mpce_shard_finish then writes a
 * NaN loss, evk_l2norm_bwd_parts NaN gradients, evk_shard_prologue stops writing into peer memory, and
 * error_host (may be NULL; an int in pinned, device-accessible HOST memory) is set as well, so the host can poll
 * it before its next step and raise. */
EVK_API int evk_peer_barrier(const uint64_t* flag_ptrs, const uint64_t* error_ptrs, int n_ranks, int rank,
                     uint32_t* epoch, int* error_host, int64_t timeout_ms, evk_stream_t stream);

/* Closes the sharded forward after the statistics exchange.  slots: [n_slots, ld_slot] fp32, slot r = rank r's
 * partial column exp-sums over its own rows (n_cols floats) followed by its row-side loss term
 * inv_count * sum_{i in rows of r} (shift + ln R_i - 2 pos_i / c_i) at index n_cols.
 *   b_col[j] = 1 / sum_r slots[r][j];   loss_out[0] = sum_r slots[r][n_cols] + inv_count * sum_j (shift + ln C_j)
 * (:501-503 on the concatenated batch).  Fixed summation order: every rank computes identical bits.
 * workspace: evk_shard_finish_workspace_bytes(n_cols) bytes, 16-byte aligned, contents irrelevant.
 * error (may be NULL): when *error != 0 (a barrier timed out, here or on a peer) loss and b_col are NaN and
 * error_host (may be NULL, pinned host int) is set.
 * workspace_persistent != 0: the caller keeps this workspace for these calls only and zeroed it once; the kernel
 * leaves its ticket zero again, so no memset node is enqueued.  sync (may be NULL): see evk_peer_sync_t - the
 * slots were stored by the peers. */
EVK_API int evk_mpce_shard_finish(const float* slots, int n_slots, int64_t ld_slot, int64_t n_cols, float shift,
                          double inv_count, float* b_col, float* loss_out,
                          void* workspace, int64_t workspace_bytes, int workspace_persistent,
                          const int* error, int* error_host, const evk_peer_sync_t* sync, evk_stream_t stream);

/* K4b with the reduce-scatter fused into its epilogue: the partial dKhat of this rank's row block,
 *   out_owner(j)[j % rows_per_owner, :] += alpha * sum_i W[i, j] x[i, :],   owner(j) = j / rows_per_owner,
 * accumulated with fp32 red.add straight into the owning rank's buffer (out_ptrs: HOST array of n_owners
 * peer-mapped device addresses, each [rows_per_owner, ld_out] fp32, zeroed by its owner before any rank
 * starts).  Replaces evk_mpce_bwd_gemm(transpose_w = 1) + ncclReduceScatter.
 * store != 0: no split-K and plain 128-bit stores instead of red.add - every element of the owners' buffers
 * is written exactly once, so out_ptrs[o] must be a buffer private to THIS source rank (the owner then adds the
 * per-source buffers up: evk_l2norm_bwd_parts) and needs no zero fill.  Posted stores use NVLink far better
 * than 16-byte atomics.  store == 2: the partials are stored as bf16 ([rows_per_owner, ld_out] bf16, ld_out % 8 == 0):
 * half the NVLink bytes; the owner still adds them up in fp32.
 * first_owner: the tiles of this owner's rows are computed and sent first, then first_owner + 1, ... (wrapping).
 * Every rank should pass a different value ((rank + 1) % n_owners): with the same order on every rank all GPUs
 * would store into the same owner at the same time and its NVLink ingress (not the sum over GPUs) would bound
 * the exchange.
 * cta_limit (also on evk_mpce_bwd_gemm; 0 = every SM): launch at most this many CTAs, so that two contractions
 * enqueued on different streams run SIDE BY SIDE on disjoint SMs (persistent CTAs use a whole SM each): in the
 * sharded backward the NVLink-bound scattering contraction and the local one overlap instead of queueing. */
EVK_API int evk_mpce_bwd_gemm_scatter(const void* w_hi, const void* w_lo, int64_t ld_w,
                              int64_t n_rows, int64_t n_cols,
                              const void* x_hi, const void* x_lo, int64_t ld_x, int64_t d,
                              float alpha, int flags,
                              const uint64_t* out_ptrs, int n_owners, int64_t rows_per_owner, int64_t ld_out,
                              int store, int first_owner, int cta_limit, evk_stream_t stream);

/* ---- f4: exact inner-product top-k retrieval ------------------------------------------------------------------
 * Replaces the faiss index of PretrainTester.predict (modules/multiview/trainer.py:543-653: IndexIVFFlat with
 * METRIC_INNER_PRODUCT, `train_index.search(ret, k)`; faiss is a third-party dependency that is not vendored in
 * the reference) by an EXACT search: scores = Q C^T on the tcgen05 main loop, folded chunk by chunk into the
 * running top-k of every query.
 * evk_tc_gemm_nt: c[m, n] = a[m, k] b[n, k]^T, bf16 operands (row-major, K contiguous, ld % 8 == 0, 16-byte aligned),
 * fp32 output written with plain stores (no zero fill needed).  a_lo / b_lo (both or neither): the low halves of a
 * (hi, lo) bf16 split of fp32 features -> 3-segment product (hi.hi + hi.lo + lo.hi, ~2^-17 relative). */
EVK_API int evk_tc_gemm_nt(const void* a_hi, const void* a_lo, int64_t lda, const void* b_hi, const void* b_lo, int64_t ldb,
                   int64_t m, int64_t n, int64_t k, float* c, int64_t ldc, evk_stream_t stream);
/* Fold a chunk of scores [n_q, n_c] (row pitch ld; column c is corpus entry col_offset + c) into the running top-k
 * lists best_val / best_idx [n_q, k] (k <= 64; sorted by score descending, ties by corpus index ascending;
 * init != 0: start from empty lists, else continue from their contents; unused slots hold -inf / -1).
 * q_group [n_q] / c_group [corpus] (both or neither): entries of the query's own group are not candidates (the
 * reference drops the retrieved reports of the query's own study, trainer.py:590-607).  NaN scores are skipped. */
EVK_API int evk_topk_update(const float* scores, int64_t ld, int64_t n_q, int64_t n_c, int64_t col_offset,
                    const int32_t* q_group, const int32_t* c_group, int k, float* best_val, int32_t* best_idx,
                    int init, evk_stream_t stream);

/* Debug/bring-up: plain C[m,n] = A[m,k] B[n,k]^T (or MN-major operands) through the same
 * tcgen05 main loop, fp32 out.  a_major/b_major: 0 = K contiguous, 1 = M/N contiguous
 * (then A is stored [k, m] / B is stored [k, n]).  c must be zeroed by the caller (the epilogue
 * accumulates with red.add).  variant selects an alternative MN-major descriptor encoding
 * (bring-up only; 0 is the production encoding); splits > 0 forces the split-K factor. */
EVK_API int evk_tc_gemm_probe(const void* a, int64_t lda, int a_major, const void* b, int64_t ldb, int b_major,
                      int64_t m, int64_t n, int64_t k, float* c, int64_t ldc, int variant, int splits,
                      evk_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* EVOKE_B200_H_ */
