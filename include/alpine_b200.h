/* alpine_b200 -- C ABI of the B200-native multiplicative-update NMF path of ALPINE.
 *
 * This is the drop-in boundary: the reference has no FFI of its own (it is pure
 * Python over torch), so each entry point below cites the reference code it replaces.  A host binding
 * (ctypes, see INTEGRATION.md and alpine_b200/_native.py) passes raw device pointers obtained from
 * torch tensors; the library never owns X, Y, W, H or B.  Every function returns 0 on success and a
 * non-zero status otherwise; alpine_last_error() returns a thread-local message for the last failure.
 * No C++ types, exceptions or torch types cross this boundary.
 *
 * Layout conventions (all float32):
 *   X  cells-major:  X[j * ldX + g], j < n cells, g < G genes, ldX % 4 == 0, 16-byte aligned base
 *      (this is what the reference holds on the device: main.py:104, 445 keep adata.X's memory order)
 *   W  G x K row-major with leading dimension ldW (ldW % 4 == 0); column blocks in covariate order,
 *      unguided block last (main.py:79)
 *   H  K x n row-major with leading dimension ldH (ldH % 4 == 0)
 *   Y_i  c_i x n row-major one-hot, all-zero column for a missing label (encoder.py:32-37, main.py:447)
 *   B_i  c_i x k_i row-major contiguous
 */
#ifndef ALPINE_B200_H
#define ALPINE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct alpine_ctx alpine_ctx;

enum { ALPINE_LOSS_KL = 0, ALPINE_LOSS_FROBENIUS = 1 };
enum { ALPINE_OK = 0, ALPINE_ERR_ARG = 1, ALPINE_ERR_CUDA = 2, ALPINE_ERR_KERNEL = 3, ALPINE_ERR_STATE = 4 };

/* ABI version of this header (bumped on any signature change). */
int alpine_abi_version(void);
/* Message of the last failing call on this thread. */
const char* alpine_last_error(void);
/* Number of kernels this library has launched in this process (bench.py's gpu_launches). */
long long alpine_launch_count(void);

/* Create a solver context for one device-resident shard of cells.
 * Replaces the per-fit setup the reference does implicitly in ALPINE.__init__/_fit (main.py:62-80, 486-496).
 *   n_blocks  = n_cov + 1; k_blocks[i] = components of block i (covariates first, unguided last)
 *   c_cov[i]  = number of categories of covariate i
 *   loss_type = ALPINE_LOSS_KL | ALPINE_LOSS_FROBENIUS (main.py:57, 371)                                   */
int alpine_create(alpine_ctx** out, int device, int64_t n_genes, int64_t n_cells, int n_blocks,
                  const int* k_blocks, int n_cov, const int* c_cov, int loss_type);
int alpine_destroy(alpine_ctx* ctx);

/* Optional: lend the context one device buffer for all of its per-fit workspaces (W^T copy, split operands, W^T X,
 * partial sums ...).  With it, creating and destroying a context performs no cudaMalloc / cudaFree (the host
 * binding takes the buffer from torch's caching allocator); without it, or for what does not fit, the library
 * allocates by itself.  Call right after alpine_create; `base` 256-byte aligned, alive until alpine_destroy.   */
int64_t alpine_workspace_bytes(const alpine_ctx* ctx);
int alpine_bind_workspace(alpine_ctx* ctx, void* base, int64_t bytes);
/* Host -> device copy of a dense fp32 matrix held in PAGEABLE host memory (what `torch.tensor(X_array, device=...)`
 * does in _initialize_matrices, main.py:445): dst[r * ld_dst + c] = src[r * ld_src + c].  `threads` host threads
 * stage row chunks through a process-wide ring of pinned buffers and queue one DMA per chunk; on return all copies
 * are queued and `stream` has been made to wait for them (no host-side wait for the GPU).  Context-free: can run
 * before alpine_create.                                                                                          */
int alpine_upload_rows(int device, float* dst, int64_t ld_dst, const float* src, int64_t ld_src, int64_t rows,
                       int64_t cols, int threads, void* stream);
/* Bind the expression matrix (AlpineMatrices.X, main.py:445). */
int alpine_bind_dense(alpine_ctx* ctx, const float* X_cells_major, int64_t ldX);
/* Bind a SPARSE expression matrix instead (north_star's optional CSR variant; the reference itself rejects sparse
 * input, main.py:395-396): device-resident CSR over cells, the layout AnnData keeps for count matrices --
 * indptr[n_cells + 1] (int64), indices[nnz] (int32 gene ids, unique within a row), values[nnz] (fp32, finite, >= 0).
 * The call converts the matrix once into the two tile-list copies the contraction kernel streams (8 bytes per
 * nonzero each, csrc/csr_tiles.cuh), computes ||X||_F^2 and synchronises `stream`; the CSR arrays are not
 * referenced afterwards.  Mathematically identical to alpine_bind_dense on the densified matrix.                */
int alpine_bind_csr(alpine_ctx* ctx, const int64_t* indptr, const int32_t* indices, const float* values, int64_t nnz,
                    void* stream);
/* Bind the one-hot label matrix of covariate i (AlpineMatrices.Ys[i], main.py:446-449). */
int alpine_bind_labels(alpine_ctx* ctx, int i, const float* Y);
/* Bind the factor matrices that the updates mutate in place (AlpineMatrices.Ws/Hs/Bs, main.py:454-470). */
int alpine_bind_factors(alpine_ctx* ctx, float* W, int64_t ldW, float* H, int64_t ldH, float* const* Bs);
/* Hyper-parameters read by the loop (main.py:64-67, 72): lam[n_cov], alpha_W, l1_ratio_W, orth_W, eps. */
int alpine_set_hparams(alpine_ctx* ctx, const double* lam, double alpha_W, double l1_ratio_W, double orth_W,
                       double eps);

/* The per-iteration all-reduce payload: [ (X H^T)^T | H H^T | rowsum(H) | per-covariate B statistics ].
 * The caller owns the buffer (so that torch.distributed can all-reduce it); size in floats.               */
int64_t alpine_reduce_buffer_size(const alpine_ctx* ctx);
int alpine_bind_reduce_buffer(alpine_ctx* ctx, float* buf);

/* Start a fit: ||X||_F^2, operand split (tf32 hi + bf16 images) of the initial H, initial statistics (H H^T, B statistics).
 * max_iter sizes the device-side loss history.  Replaces the head of _fit (main.py:486-498).              */
int alpine_fit_begin(alpine_ctx* ctx, int max_iter, void* stream);
/* The cells of one mini-batch into this context's bound arrays, one launch (replaces the advanced-indexing gathers
 * X[idx], Hs[j][:, idx], Ys[i][:, idx] of main.py:593-595): rows idx[0..cnt) of the cells-major X_all (n_all x n_genes,
 * pitch ldX_all) -> X_batch (the array bound with alpine_bind_dense; pass NULL / NULL for a CSR context, whose batch
 * rows the caller binds with alpine_bind_csr), columns idx of H_all (K x n_all, pitch ldH_all) -> the bound H, columns idx of
 * every Y_all[i] (c_i x n_all, contiguous) -> Y_batch[i] (the arrays bound with alpine_bind_labels).  Cells [cnt, n_cells)
 * of the context are zero-filled (an all-zero cell contributes nothing and stays zero), so one context serves every
 * batch of at most n_cells cells.  idx: device array of int64 cell numbers in [0, n_all); a number outside raises
 * ALPINE_ERR_ARG at the next alpine_fit_losses.                                                                   */
int alpine_batch_gather(alpine_ctx* ctx, const float* X_all, int64_t ldX_all, float* X_batch, const float* H_all,
                        int64_t ldH_all, const float* const* Y_all, float* const* Y_batch, int64_t n_all,
                        const int64_t* idx, int64_t cnt, void* stream);
/* H_all[:, idx[j]] = H[:, j] for j < cnt after the step (main.py:662).  Duplicate cell numbers carry identical columns. */
int alpine_batch_scatter(alpine_ctx* ctx, float* H_all, int64_t ldH_all, int64_t n_all, const int64_t* idx, int64_t cnt,
                         void* stream);
/* Start a mini-batch step on freshly gathered batch data (main.py:509-521): refreshes the W^T copy from the bound W
 * (another context may have updated it; skipped when this context's own W^T is still ahead of the bound W, i.e. no
 * alpine_sync_w since its last update) and the statistics of the bound H / B, without the ||X||^2 pass.  Follow
 * with alpine_mu_partials + alpine_mu_apply(ctx, 0, ...); the loss terms of a batch step are not meaningful.     */
int alpine_batch_begin(alpine_ctx* ctx, void* stream);
/* First half of one full-batch iteration: numerator X H^T of the W update (main.py:596) into the reduce
 * buffer, next to the statistics written by the previous iteration.  No data-path collective inside.      */
int alpine_mu_partials(alpine_ctx* ctx, void* stream);
/* Second half, after the (optional) all-reduce of the reduce buffer: W update (main.py:597-612), B updates
 * (main.py:615-628), H update (main.py:631-663), loss terms of iteration `iter` (main.py:666, 726-753) and
 * the statistics for the next iteration.  Nine kernel launches per iteration on one GPU together with
 * alpine_mu_partials: the two sweeps of X (W^T W rides along with W^T X), three small tcgen05 plans for the K-deep
 * products, the two fused update kernels and their two finish kernels (csrc/mu_update_kernels.cuh).                                                                           */
int alpine_mu_apply(alpine_ctx* ctx, int iter, void* stream);
/* The updates keep W as W^T inside the context and refresh the caller's row-major W only on demand: by
 * alpine_fit_losses, alpine_scale, alpine_transform, the next alpine_fit_begin / alpine_batch_begin -- or explicitly
 * with alpine_sync_w when the caller wants to read W (or hand it to another context) between iterations.        */
int alpine_sync_w(alpine_ctx* ctx, void* stream);
/* Cell sharding over NVLink peer memory instead of a library collective (csrc/peer_exchange.cuh).  One process per
 * GPU: every rank calls alpine_peer_export right after alpine_create (it allocates the rank's exchange block --
 * reduce buffer + W^T + flags -- and returns its 64-byte CUDA IPC handle), the handles are exchanged by the host
 * (any transport), every rank calls alpine_peer_import with all `world` handles (world <= 8), and a host barrier
 * follows.  Then one iteration is  alpine_mu_partials + alpine_mu_apply_peer : the W update runs on this rank's
 * gene slice with the numerator summed straight from the peers' memory and the new slice stored into every peer's
 * W^T; W and B stay bit-identical across ranks.  Destroy the contexts only after a host barrier.               */
int alpine_peer_export(alpine_ctx* ctx, void* ipc_handle_64_bytes);
int alpine_peer_import(alpine_ctx* ctx, int rank, int world, const void* ipc_handles);
int alpine_mu_apply_peer(alpine_ctx* ctx, int iter, void* stream);
/* Return to the caller-side all-reduce (alpine_mu_apply), e.g. when some other rank failed to map the peers. */
int alpine_peer_disable(alpine_ctx* ctx);
/* Block Gauss-Seidel ("ALS") sweep, use_als=True (main.py:523-588).  One iteration is
 *   alpine_mu_partials            X H^T for all blocks in one sweep of X (every H_b is still unchanged when its W_b
 *                                 update consumes its columns)           [+ all-reduce of the whole reduce buffer]
 *   for b in 0 .. n_blocks-1:     alpine_als_block(ctx, b): W_b, B_b, H_b updates (main.py:527-588), one k_b-wide
 *                                 sweep of X for W_b^T X, then this shard's new H H^T into the reduce buffer
 *                                 [+ all-reduce of its K*K floats at alpine_reduce_stats_offset, except after the last]
 *   alpine_als_finish(ctx, iter)  row-major W, loss terms of the iteration, statistics for the next one.          */
int64_t alpine_reduce_stats_offset(const alpine_ctx* ctx);
int alpine_als_block(alpine_ctx* ctx, int block, void* stream);
int alpine_als_finish(alpine_ctx* ctx, int iter, void* stream);
/* Synchronise and read back the loss terms: xnorm2 = ||X||_F^2 and, per iteration, 2 + n_cov doubles
 * [ tr(W^T X H^T), tr(W^T W H H^T), pred_0, ... ] of THIS shard; all are additive across shards and
 * recon = xnorm2 - 2*t1 + t2 (the trace identity that replaces main.py:736).  Also reports kernel faults. */
int alpine_fit_losses(alpine_ctx* ctx, int n_iter, double* xnorm2, double* rows, void* stream);

/* The loss terms of the factors AS THEY ARE (_compute_loss, main.py:726-753, without materialising W H): terms[0] =
 * tr(W^T X H^T), terms[1] = tr(W^T W H H^T), terms[2 + i] = prediction loss of covariate i, all of THIS shard and
 * additive across shards; recon = ||X||^2 - 2 terms[0] + terms[1] with ||X||^2 from alpine_fit_losses.  One sweep of X
 * on the tcgen05 kernel, fp64 accumulation; synchronises the stream.  Used for the per-epoch loss of mini-batch fits. */
int alpine_eval_loss(alpine_ctx* ctx, double* terms, void* stream);

/* Post-fit scaling (_scale_matrices, main.py:772-781). */
int alpine_scale(alpine_ctx* ctx, void* stream);

/* H-only transform loop (_transform, main.py:705-709) on the bound X and H with the bound (fixed) W:
 *   A = W^T X once, T = W^T W once, then n_iter times  H *= 2A / max(2 T H, eps).                        */
int alpine_transform(alpine_ctx* ctx, int n_iter, void* stream);

/* The two dense contractions on their own (used by tests and for roofline measurement):
 *   alpine_xh_product: out[k * ld_out + g] = sum_j X[g][j] * H[k][j]      (main.py:596 without the factor 2)
 *   alpine_wx_product: out[k * ld_out + j] = sum_g W[g][k] * X[g][j]      (main.py:653 without the factor 2)
 * using the bound X and the bound H / W.  ld_out % 4 == 0.                                                */
int alpine_xh_product(alpine_ctx* ctx, float* out, int64_t ld_out, void* stream);
int alpine_wx_product(alpine_ctx* ctx, float* out, int64_t ld_out, void* stream);

/* Device-side timing of the contraction kernel (CUDA events on the launch stream around every mu_gemm launch):
 * alpine_profile(ctx, 1) starts / resets, alpine_profile(ctx, 0) stops; alpine_profile_read synchronises the
 * recorded events and returns the summed kernel time and the number of launches.  bench.py's roofline figure.   */
int alpine_profile(alpine_ctx* ctx, int enable);
int alpine_profile_read(alpine_ctx* ctx, double* gemm_ms_total, long long* gemm_launches);

/* Introspection for bench.py: SM count, grid and pipeline depth used by the contraction kernels. */
int alpine_query(const alpine_ctx* ctx, int* num_sms, int* gemm_grid, int* smem_stages, int* k_padded);

#ifdef __cplusplus
}
#endif
#endif /* ALPINE_B200_H */
