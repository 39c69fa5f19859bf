/*
 * buffalo_b200.h -- C ABI of the B200-native matrix-factorisation training backend.
 *
 * This is the drop-in boundary: every entry point below replaces one method of the
 * reference's Cython holder classes (the `self.obj` object driven by
 * buffalo/algo/{als,bpr,warp}.py).  Plain pointers and sizes only -- no torch, numpy or
 * C++ types cross the boundary.  INTEGRATION.md shows the ctypes / Cython stub a
 * reference maintainer would add to bind it.  All file:line citations are relative to
 * the reference repository root.
 *
 * Conventions
 *  - every function returning `int` returns 0 on success and a non-zero code on failure;
 *    bfl_last_error() then returns a thread-local human-readable message.  (The reference
 *    throws std::runtime_error through CHECK_CUDA, include/buffalo/cuda/utils.cuh:24-31;
 *    `init` returns false on an unreadable/invalid option file, lib/algo.cc:22-34.)
 *  - factor matrices are float32 row-major with row pitch `vdim` = bfl_*_get_vdim()
 *    (the reference pads to a multiple of 32, lib/cuda/als/als.cu:251-252; we pad to a
 *    multiple of 4 so rows are 16-byte aligned); padding columns must be zero.
 *  - CSR layout contract (buffalo/data/base.py:187-192): `indptr[x]` is the EXCLUSIVE END
 *    offset of row x (no leading zero), int64; `keys` int32 zero-based opposite index;
 *    `vals` float32.
 *  - "host" entry points take host pointers and perform the H2D/D2H copies themselves,
 *    exactly like the reference CUDA backend (als.cu:361-364,403); "device" entry points
 *    take device pointers (e.g. torch CUDA tensor storage) and run entirely on `stream`.
 *  - there is no CPU fallback: every entry point fails loudly when no sm_100 device is
 *    present.
 */
#ifndef BUFFALO_B200_H_
#define BUFFALO_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BFL_OK 0
#define BFL_ERR_OPTION 1   /* option file missing / not parseable / unsupported value */
#define BFL_ERR_CUDA 2     /* a CUDA runtime call failed */
#define BFL_ERR_STATE 3    /* call sequence violated (e.g. update before initialize_model) */
#define BFL_ERR_ARG 4      /* bad argument */

const char* bfl_last_error(void);
/* library/ABI version and the SM architecture the kernels were compiled for (100) */
int bfl_abi_version(void);
int bfl_compiled_sm(void);
/* number of kernels this library launched since load (bench.py's gpu_launches claim) */
int64_t bfl_kernel_launch_count(void);
/* Map another process's device allocation into this process with the CURRENT device as accessor
 * (cudaIpcOpenMemHandle + lazy peer access): `handle64` is the 64-byte cudaIpcMemHandle_t exported by the owner.
 * Returns the base address of the allocation (NULL on failure, see bfl_last_error).  Used by the fused multi-GPU
 * exchange; a handle must be opened at most once per process. */
void* bfl_ipc_open(const void* handle64);
int bfl_ipc_close(void* base);
/* Device allocations that can be exported to other processes (plain cudaMalloc, so the IPC handle refers to
 * exactly this buffer): used for the factor replicas of the fused multi-GPU exchange. */
void* bfl_dev_alloc(size_t bytes);
int bfl_dev_free(void* p);
int bfl_ipc_export(void* dev_ptr, void* out_handle64);

/* ======================================================================================
 * ALS  -- replaces CyALS (buffalo/algo/_als.pyx:28-63 -> als::CALS, lib/algo_impl/als/als.cc)
 *         and the CUDA holder (buffalo/algo/cuda/_als.pyx:25-67 -> cuda_als::CuALS,
 *         lib/cuda/als/als.cu)
 * ====================================================================================== */
typedef struct bfl_als bfl_als_t;

/* CyALS.__cinit__ / __dealloc__ (_als.pyx:32-37) */
bfl_als_t* bfl_als_create(void);
void bfl_als_destroy(bfl_als_t* h);

/* CALS::init(opt_path) (als.cc:30-69; CuALS::init als.cu:230-267).  `opt_path` is the JSON
 * option file the Python layer writes (buffalo/algo/base.py:18-24).  Applies the
 * d >= 128 => "ialspp" rule (als.cc:46).  Optimizers: llt, ldlt, manual_cg, ialspp;
 * the Eigen iterative solvers (eigen_cg...eigen_minres, lib/algo.cc:83-127) are rejected
 * with BFL_ERR_OPTION. */
int bfl_als_init(bfl_als_t* h, const char* opt_path);
/* same, from JSON text already in memory */
int bfl_als_init_json(bfl_als_t* h, const char* json_text);

/* CuALS::get_vdim (als.cu:338-340) */
int bfl_als_get_vdim(bfl_als_t* h);

/* CALS::initialize_model(P, P_rows, Q, Q_rows) (als.cc:76-83; CuALS als.cu:269-289).
 * HOST pointers, [rows x vdim] float32.  The pointers are retained (reference semantics):
 * bfl_als_partial_update writes the updated rows back into them. */
int bfl_als_initialize_model(bfl_als_t* h, float* P, int32_t P_rows, float* Q, int32_t Q_rows);

/* CuALS::set_placeholder(lindptr, rindptr, batch_size) (als.cu:291-307): copies both
 * end-offset arrays to the device and sizes the key/value staging buffers. */
int bfl_als_set_placeholder(bfl_als_t* h, const int64_t* lindptr, const int64_t* rindptr, size_t batch_size);

/* CALS::precompute(axis) (als.cc:86-93; als.cu:310-319): FF = Y^T Y of the opposite
 * factor matrix (axis 0 -> Q^T Q). */
int bfl_als_precompute(bfl_als_t* h, int axis);

/* CALS::partial_update(start_x, next_x, indptr, keys, vals, axis) -> pair<double,double>
 * (als.cc:95-105 -> _partial_update :107-209 | _partial_update_ialspp :211-358;
 * CuALS::partial_update als.cu:342-406).  HOST buffers: `indptr` is the global end-offset
 * array, `keys`/`vals` are the chunk buffers starting at row start_x.  Copies the chunk to
 * the device, solves rows [start_x, next_x), copies the updated rows back into the host
 * factor matrix given to initialize_model, returns the loss numerator / denominator. */
int bfl_als_partial_update(bfl_als_t* h, int32_t start_x, int32_t next_x, const int64_t* indptr,
                           const int32_t* keys, const float* vals, int axis,
                           double* loss_nume, double* loss_deno);

/* ---- device-resident path (no reference counterpart: the reference re-uploads every
 * chunk, als.cu:361-364).  Pointers are DEVICE pointers owned by the caller. ---- */
int bfl_als_bind_factors_device(bfl_als_t* h, float* dP, int64_t P_rows, float* dQ, int64_t Q_rows);
/* bind one CSR orientation (axis 0: rowwise/users, 1: colwise/items) resident on device */
int bfl_als_bind_csr_device(bfl_als_t* h, int axis, const int64_t* d_indptr, const int32_t* d_keys,
                            const float* d_vals, int64_t rows, int64_t nnz);
/* FF = Y^T Y on `stream` */
int bfl_als_precompute_device(bfl_als_t* h, int axis, void* stream);
/* partial FF over rows [row_begin,row_end) of the opposite factor (axis 0: rows of Q) -- a row-sharded run computes
 * the Gram of its own freshly solved rows and all-reduces the d x d result (bfl_als_gram_device_mut) instead of every
 * rank recomputing the full matrix (als.cc:86-93 restricted to a row range). */
int bfl_als_precompute_rows_device(bfl_als_t* h, int axis, int64_t row_begin, int64_t row_end, void* stream);
/* solve rows [row_begin, row_end) of the bound CSR `axis` on `stream`; adds the loss pieces
 * into d_loss[0] (numerator), d_loss[1] (denominator) (device doubles, may be NULL). */
int bfl_als_update_device(bfl_als_t* h, int axis, int64_t row_begin, int64_t row_end,
                          double* d_loss, void* stream);
/* multi-GPU fused exchange (no reference counterpart; the reference is single-device): DEVICE pointers, valid in
 * this process (CUDA IPC / peer access), of the OTHER ranks' replicas of the matrix updated on `axis` (P for axis
 * 0, Q for axis 1).  Every solved row is then also stored into those replicas from inside the solve kernel, so the
 * all-gather of the updated shard overlaps the solve row by row; the caller only needs a stream-ordered barrier
 * between half-epochs.  n_peers = 0 switches the fused exchange off.  At most 15 peers. */
int bfl_als_set_peer_replicas(bfl_als_t* h, int axis, int n_peers, float* const* peer_ptrs);
/* device pointer of the current Gram matrix [d x d] (tests) */
const float* bfl_als_gram_device(bfl_als_t* h);
/* multi-GPU: when several ranks each computed the Gram of their shard of Y, the host
 * all-reduces this buffer (d*d floats) before bfl_als_update_device. */
float* bfl_als_gram_device_mut(bfl_als_t* h);

/* ======================================================================================
 * BPRMF / WARP -- replaces CyBPRMF / CyWARP (buffalo/algo/_bpr.pyx:34-92, _warp.pyx:34-92 ->
 * bpr::CBPRMF lib/algo_impl/bpr/bpr.cc, warp::CWARP lib/algo_impl/warp/warp.cc, both on
 * SGDAlgorithm lib/algo.cc:133-492) and the CUDA holder CyBPR (buffalo/algo/cuda/_bpr.pyx:27-80
 * -> cuda_bpr::CuBPR lib/cuda/bpr/bpr.cu).  The reference has no CUDA WARP (warp.py:30-32).
 * ====================================================================================== */
typedef struct bfl_sgd bfl_sgd_t;

#define BFL_SGD_BPR 0
#define BFL_SGD_WARP 1

bfl_sgd_t* bfl_sgd_create(int kind);
void bfl_sgd_destroy(bfl_sgd_t* h);

/* CBPRMF::init / CWARP::init (bpr.cc:39-47, warp.cc:71-88) */
int bfl_sgd_init(bfl_sgd_t* h, const char* opt_path);
int bfl_sgd_init_json(bfl_sgd_t* h, const char* json_text);
int bfl_sgd_get_vdim(bfl_sgd_t* h);

/* SGDAlgorithm::initialize_model(P, P_rows, Q, Q_rows, Qb, num_total_samples)
 * (algo.cc:148-176; CuBPR bpr.cu:284-312).  HOST pointers; retained; synchronised back by
 * bfl_sgd_synchronize(h, 1) / bfl_sgd_update_parameters.  Allocates gradient / momentum /
 * velocity state unless optimizer == "sgd" (algo.cc:221-254). */
int bfl_sgd_initialize_model(bfl_sgd_t* h, float* P, int32_t P_rows, float* Q, int32_t Q_rows,
                             float* Qb, int64_t num_total_samples);
/* device-resident variant: caller-owned DEVICE pointers, nothing is copied back */
int bfl_sgd_bind_factors_device(bfl_sgd_t* h, float* dP, int64_t P_rows, float* dQ, int64_t Q_rows,
                                float* dQb, int64_t num_total_samples);

/* CBPRMF::set_cumulative_table(cum_table, size) (bpr.cc:66-70): HOST int64[size]; copied. */
int bfl_sgd_set_cumulative_table(bfl_sgd_t* h, const int64_t* cum_table, int32_t size);

/* CuBPR::set_placeholder(indptr, batch_size) (bpr.cu:314-325) */
int bfl_sgd_set_placeholder(bfl_sgd_t* h, const int64_t* indptr, size_t batch_size);
/* bind a device-resident rowwise CSR (keys only) */
int bfl_sgd_bind_csr_device(bfl_sgd_t* h, const int64_t* d_indptr, const int32_t* d_keys,
                            int64_t rows, int64_t nnz);

/* SGDAlgorithm::launch_workers / wait_until_done / join (algo.cc:211-219,467-492): the GPU
 * path is stream-ordered, so these only synchronise. */
int bfl_sgd_launch_workers(bfl_sgd_t* h);
int bfl_sgd_wait_until_done(bfl_sgd_t* h);
int bfl_sgd_join(bfl_sgd_t* h, double* out);

/* SGDAlgorithm::add_jobs(start_x, next_x, indptr, positives) (algo.cc:308-362) followed by
 * the work CBPRMF::worker / CWARP::worker would do for those rows (bpr.cc:72-188,
 * warp.cc:103-173; CuBPR::partial_update bpr.cu:350-430).  HOST buffers. */
int bfl_sgd_add_jobs(bfl_sgd_t* h, int32_t start_x, int32_t next_x, const int64_t* indptr,
                     const int32_t* keys);
/* same over rows [row_begin,row_end) of the bound device CSR, on `stream` */
int bfl_sgd_add_jobs_device(bfl_sgd_t* h, int64_t row_begin, int64_t row_end, void* stream);

/* SGDAlgorithm::update_parameters (algo.cc:382-465) + CWARP projection (warp.cc:192-201);
 * on the host-pointer path it also copies P,Q,Qb back (cuda/_bpr.pyx:60-61). */
int bfl_sgd_update_parameters(bfl_sgd_t* h);
int bfl_sgd_update_parameters_device(bfl_sgd_t* h, void* stream);
/* CuBPR::synchronize(device_to_host) (bpr.cu:327-348) */
int bfl_sgd_synchronize(bfl_sgd_t* h, int device_to_host);

/* CBPRMF::compute_loss / CWARP::compute_loss (bpr.cc:227-244, warp.cc:205-226).  HOST int32[n]. */
int bfl_sgd_compute_loss(bfl_sgd_t* h, int32_t n, const int32_t* users, const int32_t* positives,
                         const int32_t* negatives, double* out_loss);

/* ---- test hooks (deterministic parity): explicit triples, gradient read-back ---- */
/* apply the BPR update to explicit DEVICE triples (what add_jobs does after sampling) */
int bfl_sgd_apply_triples_device(bfl_sgd_t* h, const int32_t* d_users, const int32_t* d_pos,
                                 const int32_t* d_neg, int64_t n, float lr, void* stream);
/* sample BPR triples for rows [row_begin,row_end) of the bound CSR into DEVICE arrays */
int bfl_sgd_sample_device(bfl_sgd_t* h, int64_t row_begin, int64_t row_end, int32_t* d_users,
                          int32_t* d_pos, int32_t* d_neg, void* stream);
/* device pointers of the gradient accumulators (NULL for optimizer == sgd) */
float* bfl_sgd_grad_device(bfl_sgd_t* h, int which /*0 P, 1 Q, 2 Qb*/);
/* device pointers of the per-row sample counters used by per_coordinate_normalize (algo.cc:398-413; int32[rows]).
 * Together with the gradient accumulators these are what a row-sharded multi-GPU epoch all-reduces before
 * update_parameters. */
int32_t* bfl_sgd_count_device(bfl_sgd_t* h, int which /*0 P rows, 1 Q rows*/);
/* WARP: per-positive trial counts / chosen negatives of the last add_jobs (device int32[nnz]) */
int bfl_sgd_set_trace_device(bfl_sgd_t* h, int32_t* d_trials, int32_t* d_negs);
/* current epoch counter / decayed learning rate (algo.cc:284-287) */
int bfl_sgd_epoch(bfl_sgd_t* h);
double bfl_sgd_current_lr(bfl_sgd_t* h);
int bfl_sgd_read_stats(bfl_sgd_t* h, double* loss_sum, int64_t* num_updates);

/* =====================================================================================
 * Evaluation top-k; replaces the host quickselect behind Evaluable.get_topk /
 * Algo._get_topk_recommendation (buffalo/evaluate/base.py:31-42, buffalo/parallel/_core.hpp:69-142):
 * scores = queries . items^T (+ item_bias), the k best item indices per query, best first; ties go to the
 * smaller index; -1 pads when there are fewer than k items.  k <= 4096.
 * ===================================================================================== */
/* device pointers, stream-ordered */
int bfl_topk_device(const float* d_queries, int64_t nq, int ldq, const float* d_items, int64_t n_items, int ldi,
                    const float* d_item_bias /* nullable */, int d, int k, int32_t* d_out_idx, float* d_out_val,
                    void* stream);
/* host pointers (copies in, runs, copies out); out_val may be NULL */
int bfl_topk_host(const float* queries, int64_t nq, int ldq, const float* items, int64_t n_items, int ldi,
                  const float* item_bias /* nullable */, int d, int k, int32_t* out_idx, float* out_val);

/* =====================================================================================
 * Ingest helpers.
 * CSR of one orientation from (major, minor, value) triples, the sort/compress stage of
 * MatrixMarket.create -> _sort_and_compressed_binarization (buffalo/data/mm.py:236-279,
 * buffalo/data/fileio.hpp:263-419): stable sort by (major, minor) (sort_minor = 0: by major only,
 * Stream's token order), indptr[num_major] = exclusive END offsets (buffalo/data/base.py:187-192).
 * Cumulative popularity table of BPRMF.prepare_sampling (buffalo/algo/bpr.py:99-111):
 * cum[i] = sum_{j<=i} count(j)^power.
 * ===================================================================================== */
int bfl_csr_from_triples_device(const int32_t* d_major, const int32_t* d_minor, const float* d_vals, int64_t nnz,
                                int32_t num_major, int32_t num_minor, int sort_minor, int64_t* d_indptr,
                                int32_t* d_key_out, float* d_val_out, void* stream);
int bfl_csr_from_triples_host(const int32_t* major, const int32_t* minor, const float* vals, int64_t nnz,
                              int32_t num_major, int32_t num_minor, int sort_minor, int64_t* indptr,
                              int32_t* key_out, float* val_out);
int bfl_popularity_table_device(const int32_t* d_keys, int64_t nnz, int32_t n_items, int power, int64_t* d_cum,
                                void* stream);
int bfl_popularity_table_host(const int32_t* keys, int64_t nnz, int32_t n_items, int power, int64_t* cum);

#ifdef __cplusplus
}
#endif
#endif /* BUFFALO_B200_H_ */
