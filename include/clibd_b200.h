/* clibd_b200 -- C ABI of the B200-native CLIBD hot path (sm_100a).
 *
 * The reference (bioscan-ml/clibd) is pure Python and has no FFI; these entry points
 * are what a binding for its hot path would call.  Each one names the reference
 * interface it replaces.  Conventions:
 *   - every pointer is a DEVICE pointer unless its comment says "host";
 *   - the caller owns all memory (inputs, outputs, scratch); the library never
 *     allocates, frees or retains device memory past the call;
 *   - all work is enqueued asynchronously on `stream`;
 *   - return 0 = ok, non-zero = error; text via clibd_last_error() (thread-local);
 *   - dtype codes: 0 = float32, 1 = bfloat16, 2 = float16;
 *   - path codes : 0 = CUDA-core fp32 (exact, any N/d),
 *                  1 = tcgen05 tensor cores, bf16 operands, fp32 accumulate,
 *                  2 = tcgen05 tensor cores, f16 operands, fp32 accumulate.
 *   - feature slots are always [image, dna, text]; an absent modality is NULL;
 *   - mode codes (row-sharded loss, n_local < n_global; ignored on one GPU):
 *                  0 = every rank produces both gradients of its rows itself (S recomputed in two sweeps per pair),
 *                  1 = exchange: S is computed once per pair on every rank; the column-side gradient leaves the
 *                      rank as partial rows that are reduce-scattered to their owners (the backward of the
 *                      reference's torch.distributed.nn.all_gather, loss_func.py:97) -- by the caller with NCCL, or
 *                      tile by tile over NVLink into peer-mapped slot arrays (clibd_loss_backward_sweeps);
 *   - pair_weight[3] weights the unordered pairs (image,dna), (image,text), (dna,text):
 *     the loss is sum_p pair_weight[p] * [CE(S_p, T) + CE(S_p^T, T)], so the reference's
 *     "mean over the filtered ordered-pair list" (loss_func.py:69,200) is
 *     pair_weight[p] = multiplicity_p / len(loss_list).
 */
#ifndef CLIBD_B200_H
#define CLIBD_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CUstream_st* clibd_stream_t; /* == cudaStream_t */

#define CLIBD_ABI_VERSION 7

int clibd_abi_version(void);
const char* clibd_last_error(void);
/* 1 if the current device is compute capability 10.x (the tcgen05 paths need it). */
int clibd_device_supported(void);

/* ---- telemetry (new; no reference counterpart) --------------------------------------
 * Number of CUDA kernels this library has launched in this process (bench.py's gpu_launches). */
int64_t clibd_kernel_launch_count(void);
/* Optional CUDA-event timing of the tensor-core kernels, recorded on their launching stream.
 * Slots: 0 loss forward (tcgen05), 1 loss backward (tcgen05), 2 kNN screen (tcgen05), 3 kNN re-rank.
 * clibd_profile_read synchronises the recorded events, writes per-slot total milliseconds and launch
 * counts into host arrays of 8 entries each, and clears them. */
int clibd_profile_enable(int enable);
/* 1 when the loss entry points replay their launch sequence as a CUDA graph for this shape (launch-bound batches:
 * n_global * n_local <= 2^28; CLIBD_GRAPHS=0 switches it off).  Event profiling bypasses the graphs, so a benchmark
 * of such a shape takes its per-kernel timings in a separate pass. */
int clibd_graphs_active(int64_t n_global, int64_t n_local);
int clibd_profile_read(double* total_ms /* host [8] */, int64_t* launches /* host [8] */);

/* ---- contrastive loss ------------------------------------------------------------
 * Replaces ContrastiveLoss.forward / ClipLoss.forward and their autograd backward
 * (bioscanclip/model/loss_func.py:41-69, :138-201) and construct_label_metrix (:19-22).
 * Phases are split where the reference all-gathers (loss_func.py:143-157) so the host
 * can run the collectives between them; with one process call them back to back. */

/* inv_norm[i] = 1 / max(||x_i||_2, 1e-12)  (F.normalize, loss_func.py:55-56,186-187) */
int clibd_row_inv_norm(const void* x, int dtype, int64_t n, int64_t d, float* inv_norm, clibd_stream_t stream);

/* Bytes of scratch the three loss calls below need (same scratch, kept from forward
 * to backward). */
int64_t clibd_loss_scratch_bytes(int64_t n_global, int64_t n_local, int64_t d, int path, int mode);

/* Forward statistics for the local row block [row0, row0+n_local) against all n_global
 * columns.  x[m]: gathered [n_global, d] row-major features in `dtype`; inv_norm[m]:
 * [n_global]; labels: [n_global] int64.
 * Writes rowsum[p*n_global + i] for LOCAL rows i (sum_j exp(S_ij - s)), colsum[p*n_global + j]
 * for all j summed over the local rows only (all-reduce it across ranks), and
 * pos[p] = sum_{i local} sum_j T_ij * cos_ij (all-reduce it).
 * logit_scale_dev: optional DEVICE scalar (the reference's learnable `logit_scale.exp()` tensor,
 * simple_clip.py:61); when given it replaces logit_scale, is copied into the scratch and every kernel of this
 * call, of clibd_loss_forward_finish and of clibd_loss_backward reads it from there -- no host read per step.
 * The scale must be a positive finite number (host value: error otherwise; device value: NaN loss); it is NOT limited
 * from above -- the reference never clamps its learnable scale -- the kernels choose their softmax shift from it.
 * labels == NULL: clibd_loss_label_stage has already run on this scratch.
 * mode 1 (exchange) additionally writes posrow[p*n_global + i] = xhat_a[i] . sum_{j: label_j = label_i} xhat_b[j] for the
 * LOCAL rows i of every weighted pair p = (a, b) (disjoint support across ranks: all-reduce / exchange it like rowsum;
 * the backward needs it for all rows); posrow may be NULL in mode 0. */
int clibd_loss_forward_stats(const void* const x[3], int dtype, const float* const inv_norm[3],
                             const int64_t* labels, int64_t n_global, int64_t d, int64_t row0, int64_t n_local,
                             float logit_scale, const float* logit_scale_dev, const float pair_weight[3] /* host */,
                             int path, int mode, void* scratch, int64_t scratch_bytes, float* rowsum, float* colsum,
                             float* posrow, double* pos, clibd_stream_t stream);

/* The label statistics of clibd_loss_forward_stats alone (class representatives and sizes, class sort, positive
 * ranges): call it as soon as the gathered labels are there, then clibd_loss_forward_stats with labels == NULL. */
int clibd_loss_label_stage(const int64_t* labels, int64_t n_global, int64_t n_local, int64_t d, int path, int mode,
                           void* scratch, int64_t scratch_bytes, clibd_stream_t stream);

/* Loss value from complete statistics (rowsum/colsum/pos now hold GLOBAL sums for all
 * n_global rows/columns); also prepares the backward coefficients inside scratch.  The scale used is the one
 * clibd_loss_forward_stats stored in the scratch (logit_scale here is informational). */
int clibd_loss_forward_finish(int64_t n_global, int64_t n_local, int64_t d, float logit_scale,
                              const float pair_weight[3] /* host */, int path, int mode, void* scratch,
                              int64_t scratch_bytes,
                              const float* rowsum, const float* colsum, const double* pos, float* loss_out,
                              clibd_stream_t stream);

/* Backward for the local rows.  dx[m]: [n_local, d] in `dtype` (NULL to skip) receives
 * grad_feat_scale * grad_feat_scale_dev[0] * dL/dx (the device scalar is optional, NULL = 1: it lets the
 * caller pass autograd's grad_output without a host read); dscale_partial[0] receives sum over the local rows'
 * contribution to dL/d(logit_scale) for unit upstream gradient (all-reduce, then
 * multiply by the rank's own grad_output).  grad_feat_scale = sum over ranks of
 * grad_output (the reduce-scatter(SUM) convention of torch.distributed.nn.all_gather,
 * loss_func.py:97).  The scale used is the one clibd_loss_forward_stats stored in the scratch (logit_scale here is
 * informational).  With n_local == n_global (one GPU) and n_global >= 4096 the backward computes S once per
 * modality pair and passes its 16-bit coefficients to a second GEMM through a strip inside the scratch
 * (environment: CLIBD_GT_STRIP_MB bounds the strip, CLIBD_BWD_TWO_SWEEPS=1 selects the two-sweep form).
 * This one-call form is for one GPU and for mode 0; mode 1 uses the two calls below. */
int clibd_loss_backward(const void* const x[3], int dtype, const float* const inv_norm[3], int64_t n_global,
                        int64_t d, int64_t row0, int64_t n_local, float logit_scale,
                        const float pair_weight[3] /* host */, int path, void* scratch, int64_t scratch_bytes,
                        float grad_feat_scale, const float* grad_feat_scale_dev, void* const dx[3],
                        double* dscale_partial, clibd_stream_t stream);

/* Backward of a row-sharded step in mode 1, split where the reduce-scatter of the column-side gradients happens.
 *
 * clibd_loss_backward_sweeps: for every weighted pair (a, b) one row sweep over the LOCAL rows of a (S tile ->
 * coefficients -> gradient of those rows, kept in the scratch) whose 16-bit coefficient tiles also feed a GEMM that
 * yields this rank's PARTIAL gradient of ALL n_global rows of b.  posrow: [3, n_global] complete (see
 * clibd_loss_forward_stats).  Where the partial rows go:
 *   peer_red == NULL: part[m] ([n_global, d] float32, caller-owned, one per column modality m) receives them (pairs
 *     that share a column modality accumulate); the caller reduce-scatters part[m] over the ranks (NCCL) into
 *     reduced[m] = 1 slot of [n_local, d];
 *   peer_red != NULL (host array [world * 3] of device pointers, entry [q * 3 + p] = the slot array of PAIR p in rank
 *     q's memory, peer-mapped, [world, n_local, d] float32): the GEMM epilogue stores row g straight into slot `rank` of
 *     its owner q = g / n_local -- the transfer rides on the GEMM, tile by tile over NVLink.  After a barrier across the
 *     ranks the owner holds world slots per pair.
 * clibd_loss_backward_finish: reduced[m] = the received column-side partials of modality m ([reduced_slots[m], n_local, d]
 * float32, summed in slot order; NULL / 0 when m is no pair's column modality) + the row-side gradient in the scratch
 * -> target terms, normalise-backward, dx, dscale_partial as in clibd_loss_backward.  grad_feat_scale_dev may point to
 * grad_count device scalars (one upstream gradient per rank) whose SUM scales the feature gradients.
 * Pair p's column modality is dna for (image, dna) and text for (image, text) and (dna, text). */
int clibd_loss_backward_sweeps(const void* const x[3], int dtype, const float* const inv_norm[3], int64_t n_global,
                               int64_t d, int64_t row0, int64_t n_local, float logit_scale,
                               const float pair_weight[3] /* host */, int path, void* scratch, int64_t scratch_bytes,
                               const float* posrow, float* const part[3], float* const peer_red[] /* host */,
                               int rank, int world, clibd_stream_t stream);
int clibd_loss_backward_finish(const void* const x[3], int dtype, const float* const inv_norm[3], int64_t n_global,
                               int64_t d, int64_t row0, int64_t n_local, float logit_scale,
                               const float pair_weight[3] /* host */, int path, void* scratch, int64_t scratch_bytes,
                               const float* const reduced[3], const int reduced_slots[3] /* host */,
                               float grad_feat_scale, const float* grad_feat_scale_dev, int grad_count,
                               void* const dx[3], double* dscale_partial, clibd_stream_t stream);

/* ---- exchanges of the row-sharded step over peer-mapped memory ------------------------------------------------
 * One process per GPU; every rank maps every other rank's exchange buffer (symmetric allocation, e.g.
 * torch.distributed._symmetric_memory) and passes HOST arrays of DEVICE pointers, entry q = the buffer inside rank
 * q's memory.  These replace the NCCL collectives of the reference (loss_func.py:143-157: all-gather of labels and
 * features) with plain stores over NVLink; the caller separates writers from readers with a barrier across ranks.
 * world <= 16.
 *
 * clibd_shard_push_rows: the all-gather.  Copies this rank's n_local rows of every present modality (raw, `dtype`),
 * their inverse norms (computed here, as clibd_row_inv_norm) and its labels into rows [rank * n_local, (rank+1) *
 * n_local) of EVERY rank's gathered buffers: peer_x[q*3+m] = [n_global, d] in `dtype`, peer_inv[q*3+m] = [n_global]
 * float32, peer_labels[q] = [n_global] int64.  labels_local == NULL: feature rows and inverse norms only; every
 * x_local[m] == NULL: labels only (a sharded step sends the labels ahead so that the label statistics run next to the
 * push of the feature rows). */
int clibd_shard_push_rows(const void* const x_local[3], int dtype, const int64_t* labels_local, int64_t n_local,
                          int64_t d, int rank, int world, void* const peer_x[] /* host */,
                          float* const peer_inv[] /* host */, int64_t* const peer_labels[] /* host */,
                          clibd_stream_t stream);
/* clibd_shard_push_stats: the statistics all-reduce, first half.  stats = this rank's [9, n_global] float32 buffer
 * rowsum[3] | colsum[3] | posrow[3] as clibd_loss_forward_stats wrote it (its own buffer of peer_stats).  The local
 * segments of rowsum and posrow (disjoint support) are copied into the same place of every other rank's
 * peer_stats[q]; the colsum partials go to slot `rank` of peer_colslots[q] ([world, 3, n_global] float32) and pos[3]
 * to slot `rank` of peer_posslots[q] ([world, 4] float64).
 * clibd_shard_reduce_stats (after the barrier): colsum of `stats` and pos[3] = sums of the world slots in rank
 * order (bit-identical on every rank). */
int clibd_shard_push_stats(const float* stats, const double* pos, int64_t n_global, int64_t row0, int64_t n_local,
                           int rank, int world, float* const peer_stats[] /* host */,
                           float* const peer_colslots[] /* host */, double* const peer_posslots[] /* host */,
                           clibd_stream_t stream);
int clibd_shard_reduce_stats(const float* colslots, const double* posslots, int64_t n_global, int world, float* stats,
                             double* pos, clibd_stream_t stream);
/* clibd_shard_push_floats: copies `count` floats into slot `rank` (of `count` floats) of every rank's peer_slots[q]
 * ([world, count] float32) -- e.g. each rank's autograd grad_output, summed by clibd_loss_backward_finish. */
int clibd_shard_push_floats(const float* src, int64_t count, int rank, int world, float* const peer_slots[] /* host */,
                            clibd_stream_t stream);
/* clibd_shard_barrier: barrier across the ranks on `stream` -- everything the ranks enqueued before it (their stores
 * into peer memory included) is complete and visible before anything enqueued after it starts.  peer_flags[q] = rank
 * q's flag block of clibd_shard_barrier_bytes() bytes in peer-mapped memory, zeroed once before first use; barriers that
 * may be in flight at the same time (two streams) use different channels (0..3).  Replaces the NCCL collective's implicit
 * synchronisation of torch.distributed.nn.all_gather (loss_func.py:97,143) in the peer form of the sharded step; takes no
 * per-call state from the host, so it can sit inside a CUDA graph.  A peer that never arrives traps after 60 s. */
int clibd_shard_barrier(uint64_t* const peer_flags[] /* host */, int rank, int world, int channel, clibd_stream_t stream);
int64_t clibd_shard_barrier_bytes(void);

/* ---- cosine nearest-neighbour retrieval --------------------------------------------
 * Replaces make_prediction / find_closest_match's search (bioscanclip/util/util.py:
 * 521-528, 759-766): sklearn L2-normalise (float64) -> float32, faiss IndexFlatIP
 * add + search(k), with ties broken by LOWEST index. */

/* out[i,:] = float32( x[i,:] / ||x[i,:]||_2 ) with the norm and division in float64;
 * zero rows stay zero.  x dtype: 0 = float32, 3 = float64. */
int clibd_knn_normalize(const void* x, int dtype, int64_t n, int64_t d, float* out, clibd_stream_t stream);

int64_t clibd_knn_scratch_bytes(int64_t n_query, int64_t n_key, int64_t d, int k, int path);

/* Exact top-k of q32 [n_query,d] against keys32 [n_key,d] (both already normalised
 * float32).  Similarity = float64 sum over d, in index order, of the exact products;
 * order key (-sim, index).  Global index = key_offset + local row.  path 1/2 screen with
 * tcgen05 16-bit operands and re-rank candidates exactly (queries whose screen cannot be
 * proven complete are re-done exhaustively in float64); path 0 is exhaustive float64.
 * n_exhaustive (device, int32[1]) receives how many queries took the exhaustive route. */
int clibd_knn_search(const float* q32, int64_t n_query, const float* keys32, int64_t n_key, int64_t key_offset,
                     int64_t d, int k, int path, void* scratch, int64_t scratch_bytes, double* out_sims64,
                     int64_t* out_idx, int32_t* n_exhaustive, clibd_stream_t stream);

/* Merge `parts` per-shard results [parts, n_query, k] into the global top-k by (-sim, index). */
int clibd_knn_merge(const double* sims64, const int64_t* idx, int parts, int64_t n_query, int k,
                    double* out_sims64, float* out_sims32, int64_t* out_idx, clibd_stream_t stream);

/* ---- top-k accuracy ------------------------------------------------------------------
 * Replaces top_k_micro_accuracy (util.py:379-395) and top_k_macro_accuracy (:555-599)
 * on integer label ids.  idx: [n_query, kmax] global key indices; key_ids: [n_key, 4]
 * ids per level (order, family, genus, species); query_ids: [n_query, 4];
 * k_list: host array of nk values <= kmax; n_class[l]: ids at level l are in [0, n_class[l]).
 * hits:  [nk, 4, n_query] uint8 scratch; class_hit/class_cnt: [nk, 4, max_class] int32 scratch
 * (zeroed by the call).  Outputs are COUNTS so the host reproduces the reference's float
 * arithmetic exactly: micro_hits [nk,4] int64; class_hit/class_cnt hold per-class counts. */
int clibd_topk_accuracy(const int64_t* idx, int64_t n_query, int kmax, const int32_t* key_ids, int64_t n_key,
                        const int32_t* query_ids, const int32_t* k_list /* host */, int nk, int32_t max_class,
                        int64_t* micro_hits, int32_t* class_hit, int32_t* class_cnt, clibd_stream_t stream);

/* ---- seams next to the hot path (SURVEY.md section 8 f, ranks 2 and 3) ----------------------------
 * Embedding hand-off: replaces `F.normalize(output, dim=-1).cpu().tolist()` per batch followed by
 * `np.array(list)` (bioscanclip/epoch/inference_epoch.py:96-101, 108-119).  Rows of x [n, d] (dtype 0/1/2)
 * are L2-normalised (x / max(||x||_2, 1e-12), float32 arithmetic) and written as float32 into rows
 * [row_offset, row_offset + n) of the caller's device store [store_rows, store_ld] (store_ld >= d). */
int clibd_embed_append(const void* x, int dtype, int64_t n, int64_t d, float* store, int64_t store_rows,
                       int64_t store_ld, int64_t row_offset, clibd_stream_t stream);

/* BarcodeBERT head: out[b, c] = mean_t softmax_c(logits[b, t, :])  (`logits.softmax(dim=-1).mean(dim=1)`,
 * bioscanclip/model/dna_encoder.py:137) and its backward
 *   grad_logits[b,t,c] = p[b,t,c] / T * (grad_out[b,c] - sum_c' grad_out[b,c'] p[b,t,c']).
 * logits / grad_logits: [n, tokens, classes] contiguous in `dtype`; out / grad_out: [n, classes] in `dtype`.
 * The probabilities are recomputed in the backward; no [n, tokens, classes] intermediate is stored.
 * scratch: clibd_softmax_mean_scratch_bytes() bytes (0 for rows of <= 1024 16-byte-aligned values). */
int64_t clibd_softmax_mean_scratch_bytes(int64_t n, int64_t tokens, int64_t classes, int dtype);
int clibd_softmax_mean_forward(const void* logits, int dtype, int64_t n, int64_t tokens, int64_t classes, void* out,
                               void* scratch, int64_t scratch_bytes, clibd_stream_t stream);
int clibd_softmax_mean_backward(const void* logits, const void* grad_out, int dtype, int64_t n, int64_t tokens,
                                int64_t classes, void* grad_logits, clibd_stream_t stream);

/* ---- SimCLR info-NCE (SURVEY.md section 8 f, rank 4) --------------------------------------------
 * Replaces SimCLR.info_nce_loss + nn.CrossEntropyLoss (bioscanclip/util/simclr.py:64-92, 118-119) for
 * n_views = 2: z [m, d] in `dtype`, rows i and (i + m/2) mod m are the two views of one image;
 * loss = mean_i [ LSE_{j != i}(cos_ij / tau) - cos_{i,partner(i)} / tau ].  The m x m logits are never
 * materialised (same fused kernels as the contrastive loss, diagonal entries excluded).
 * inv_norm: [m] from clibd_row_inv_norm; scratch: clibd_loss_scratch_bytes(m, m, d, path) bytes, kept from
 * forward to backward; rowsum: [m] float32 output (sum_{j != i} exp((cos_ij - 1) / tau), kept for the caller's
 * diagnostics); inv_temperature = 1 / tau > 0.  tcgen05 paths need d <= 768.
 * Backward: dz [m, d] in `dtype` receives grad_scale * grad_scale_dev[0] * dL/dz (device scalar optional). */
int clibd_infonce_forward(const void* z, int dtype, const float* inv_norm, int64_t m, int64_t d,
                          float inv_temperature, int path, void* scratch, int64_t scratch_bytes, float* rowsum,
                          float* loss_out, clibd_stream_t stream);
int clibd_infonce_backward(const void* z, int dtype, const float* inv_norm, int64_t m, int64_t d,
                           float inv_temperature, int path, void* scratch, int64_t scratch_bytes, float grad_scale,
                           const float* grad_scale_dev, void* dz, clibd_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* CLIBD_B200_H */
