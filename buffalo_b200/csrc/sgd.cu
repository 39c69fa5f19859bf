// BPRMF / WARP backend: sampling, pairwise update, optimizer and loss kernels + C ABI.
// Replaces bpr::CBPRMF (lib/algo_impl/bpr/bpr.cc), warp::CWARP (lib/algo_impl/warp/warp.cc),
// SGDAlgorithm (lib/algo.cc:133-492) and cuda_bpr::CuBPR (lib/cuda/bpr/bpr.cu).
//
// Work decomposition: the reference packs CSR rows into jobs and feeds worker threads through a
// mutex queue (algo.cc:308-362, concurrent_queue.hpp); here the unit of work is one positive
// (u, pos) of the chunk, one warp each, grid-striding.  Every random draw is a pure function of
// (seed, epoch, global positive index, draw number) -- Philox4x32-10, bit-identical to
// oracle/buffalo_oracle.c -- so WARP epochs and BPR adagrad/adam epochs are reproducible and can be
// compared element-wise with the oracle; plain-SGD BPR is Hogwild (atomics) like the reference.
#include "bfl_common.cuh"

using namespace bfl;

namespace {

struct SgdArgs {
    float* P;
    float* Q;
    float* Qb;
    float* gP;
    float* gQ;
    float* gQb;
    int32_t* cP;
    int32_t* cQ;
    const int64_t* indptr;   // global end offsets (device)
    const int32_t* keys;     // element (it - shift)
    const int64_t* cum;      // cumulative popularity table or null
    int32_t* trace_trials;   // optional WARP trace (indexed it - shift)
    int32_t* trace_negs;
    double* stat_loss;       // device double
    unsigned long long* stat_updates;
    int64_t shift;
    int64_t row_begin, row_end;
    int64_t it_begin, it_end;  // global positive index range of the chunk
    int32_t num_items;
    int D, ld;
    int optimizer;  // 0 sgd 1 adagrad 2 adam
    int use_bias, update_i, update_j, num_neg, verify_neg, uniform, pcn, max_trials, score_l2;
    uint32_t seed, epoch;
    float reg_u, reg_i, reg_j, reg_b, lr, threshold;
};

// smallest row r in [lo, hi) with indptr[r] > it
__device__ __forceinline__ int64_t row_of(const int64_t* __restrict__ indptr, int64_t lo, int64_t hi, int64_t it) {
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (__ldg(indptr + mid) > it) hi = mid; else lo = mid + 1;
    }
    return lo;
}

__device__ __forceinline__ bool seen_sorted(const int32_t* __restrict__ keys, int64_t n, int32_t item) {
    int64_t lo = 0, hi = n;
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (__ldg(keys + mid) < item) lo = mid + 1; else hi = mid;
    }
    return lo < n && __ldg(keys + lo) == item;
}

__device__ __forceinline__ int32_t cum_lower_bound(const int64_t* __restrict__ cum, int32_t size, int64_t r) {
    int32_t lo = 0, hi = size;
    while (lo < hi) {
        const int32_t mid = (lo + hi) >> 1;
        if (__ldg(cum + mid) < r) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// ---------------------------------------------------------------------------------------
// BPR negative sampling (bpr.cc:106-117; reference GPU: fill_rows + generate_samples,
// bpr.cu:22-87).  One thread per (positive, k < num_neg).
// ---------------------------------------------------------------------------------------
__global__ void bpr_sample_kernel(SgdArgs a, int32_t* __restrict__ out_u, int32_t* __restrict__ out_pos,
                                  int32_t* __restrict__ out_neg) {
    const int64_t total = (a.it_end - a.it_begin) * a.num_neg;
    for (int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; s < total; s += (int64_t)gridDim.x * blockDim.x) {
        const int64_t it = a.it_begin + s / a.num_neg;
        const int k = (int)(s % a.num_neg);
        const int64_t row = row_of(a.indptr, a.row_begin, a.row_end, it);
        const int64_t beg = row == 0 ? 0 : __ldg(a.indptr + row - 1);
        const int64_t end = __ldg(a.indptr + row);
        const int32_t* rk = a.keys + (beg - a.shift);
        const uint64_t sid = (uint64_t)it * a.num_neg + k;
        int32_t neg = 0;
        for (uint32_t t = 0;; ++t) {
            if (a.uniform) {
                neg = draw_range(a.seed, a.epoch, sid, t, (uint32_t)a.num_items);
            } else {
                const uint64_t tot = (uint64_t)__ldg(a.cum + a.num_items - 1);
                const uint64_t r64 = ((uint64_t)draw_u32(a.seed, a.epoch, sid, 2 * t) << 32) |
                                     draw_u32(a.seed, a.epoch, sid, 2 * t + 1);
                const int64_t r = (int64_t)__umul64hi(r64, tot);
                neg = cum_lower_bound(a.cum, a.num_items, r);
                if (neg >= a.num_items) neg = a.num_items - 1;
            }
            if (!a.verify_neg || !seen_sorted(rk, end - beg, neg)) break;
            if (t >= 64) break;
        }
        out_u[s] = (int32_t)row;
        out_pos[s] = __ldg(a.keys + (it - a.shift));
        out_neg[s] = neg;
    }
}

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
// L2-coherent read: rows are being updated by atomics from other SMs (and by this warp's previous triple)
__device__ __forceinline__ float4 ld4_cg(const float* p) { return __ldcg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ void red4(float* p, float4 v) { atomicAdd(reinterpret_cast<float4*>(p), v); }

// ---------------------------------------------------------------------------------------
// BPR pairwise step (bpr.cc:119-171; reference GPU update_bpr_kernel bpr.cu:89-146).
// One warp per triple; rows are read/updated as float4 (row pitch ld is a multiple of 4,
// padding columns are zero and stay zero).  NV = ceil(ld / 128).
// sgd: deltas are formed from the values read before the step and applied with vector atomics
// (pre-update form, like bpr.cu:122-134).  adagrad/adam: gradients are accumulated (bpr.cc:138-156).
// ---------------------------------------------------------------------------------------
template <int NV>
__global__ void __launch_bounds__(256) bpr_apply_kernel(SgdArgs a, const int32_t* __restrict__ us,
                                                       const int32_t* __restrict__ poss,
                                                       const int32_t* __restrict__ negs, int64_t n) {
    const int lane = threadIdx.x & 31;
    const int64_t w0 = (int64_t)blockIdx.x * (blockDim.x >> 5) + warp_id_uniform();
    const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int nv4 = a.ld >> 2;
    // Each warp walks a CONTIGUOUS range of triples one after the other, like a worker thread of the reference
    // walks its job's rows (bpr.cc:103-117): the positives of one user are applied sequentially and concurrent
    // warps work on different users, so only item rows are shared Hogwild-style.
    const int64_t per = (n + nw - 1) / nw;
    const int64_t s_end = (w0 + 1) * per < n ? (w0 + 1) * per : n;
    for (int64_t s = w0 * per; s < s_end; ++s) {
        const int u = __ldg(us + s), pos = __ldg(poss + s), neg = __ldg(negs + s);
        float* pu = a.P + (int64_t)u * a.ld;
        float* qi = a.Q + (int64_t)pos * a.ld;
        float* qj = a.Q + (int64_t)neg * a.ld;
        float4 vp[NV], vi[NV], vj[NV];
        float part = 0.f;
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            const int c = lane + 32 * k;
            if (c < nv4) {
                vp[k] = ld4_cg(pu + 4 * c);
                vi[k] = ld4_cg(qi + 4 * c);
                vj[k] = ld4_cg(qj + 4 * c);
                part += vp[k].x * (vi[k].x - vj[k].x) + vp[k].y * (vi[k].y - vj[k].y) +
                        vp[k].z * (vi[k].z - vj[k].z) + vp[k].w * (vi[k].w - vj[k].w);
            }
        }
        float x = warp_sum(part);  // bpr.cc:119
        float bi = 0.f, bj = 0.f;
        if (a.use_bias) {
            bi = __ldcg(a.Qb + pos);
            bj = __ldcg(a.Qb + neg);
            x += bi - bj;  // bpr.cc:120-121
        }
        // logit = 1 - sigmoid(x) with the reference's clamp at +-MAX_EXP = 6 (bpr.cc:123-131; the exact
        // expression of bpr.cu:113-116 instead of the CPU path's 1000-entry table)
        const float logit = x > 6.f ? 0.f : (x < -6.f ? 1.f : 1.0f / (1.0f + __expf(x)));
        if (a.optimizer != 0) {
#pragma unroll
            for (int k = 0; k < NV; ++k) {
                const int c = lane + 32 * k;
                if (c < nv4) {
                    red4(a.gP + (int64_t)u * a.ld + 4 * c,
                         make_float4(logit * (vi[k].x - vj[k].x), logit * (vi[k].y - vj[k].y),
                                     logit * (vi[k].z - vj[k].z), logit * (vi[k].w - vj[k].w)));
                    const float4 g = make_float4(logit * vp[k].x, logit * vp[k].y, logit * vp[k].z, logit * vp[k].w);
                    if (a.update_i) red4(a.gQ + (int64_t)pos * a.ld + 4 * c, g);
                    if (a.update_j) red4(a.gQ + (int64_t)neg * a.ld + 4 * c, make_float4(-g.x, -g.y, -g.z, -g.w));
                }
            }
            if (lane == 0) {
                if (a.use_bias) {
                    if (a.update_i) atomicAdd(a.gQb + pos, logit);
                    if (a.update_j) atomicAdd(a.gQb + neg, -logit);
                }
                if (a.pcn) {  // bpr.cc:140-143 per sample; :174-181 once per positive
                    atomicAdd(a.cQ + neg, 1);
                    if (s % a.num_neg == 0) {
                        atomicAdd(a.cP + u, 1);
                        atomicAdd(a.cQ + pos, 1);
                    }
                }
            }
        } else {
            const float lr = a.lr;
#pragma unroll
            for (int k = 0; k < NV; ++k) {
                const int c = lane + 32 * k;
                if (c < nv4) {
                    if (a.update_i)
                        red4(qi + 4 * c, make_float4(lr * (logit * vp[k].x - a.reg_i * vi[k].x),
                                                     lr * (logit * vp[k].y - a.reg_i * vi[k].y),
                                                     lr * (logit * vp[k].z - a.reg_i * vi[k].z),
                                                     lr * (logit * vp[k].w - a.reg_i * vi[k].w)));
                    if (a.update_j)
                        red4(qj + 4 * c, make_float4(lr * (-logit * vp[k].x - a.reg_j * vj[k].x),
                                                     lr * (-logit * vp[k].y - a.reg_j * vj[k].y),
                                                     lr * (-logit * vp[k].z - a.reg_j * vj[k].z),
                                                     lr * (-logit * vp[k].w - a.reg_j * vj[k].w)));
                    red4(pu + 4 * c, make_float4(lr * (logit * (vi[k].x - vj[k].x) - a.reg_u * vp[k].x),
                                                 lr * (logit * (vi[k].y - vj[k].y) - a.reg_u * vp[k].y),
                                                 lr * (logit * (vi[k].z - vj[k].z) - a.reg_u * vp[k].z),
                                                 lr * (logit * (vi[k].w - vj[k].w) - a.reg_u * vp[k].w)));
                }
            }
            if (lane == 0 && a.use_bias) {
                if (a.update_i) atomicAdd(a.Qb + pos, lr * (logit - a.reg_b * bi));
                if (a.update_j) atomicAdd(a.Qb + neg, lr * (-logit - a.reg_b * bj));
            }
        }
    }
}

// ---------------------------------------------------------------------------------------
// WARP rank sampling + gradient accumulation (warp.cc:103-173).  One warp per positive.
// P and Q are read-only inside an epoch (gradients only, warp.cc:156-158).
// ---------------------------------------------------------------------------------------
template <int NV>
__device__ __forceinline__ float warp_score(const float4 (&vp)[NV], const float* __restrict__ q, int nv4, int lane,
                                            int l2, float4 (&vq)[NV]) {
    float part = 0.f;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        const int c = lane + 32 * k;
        if (c < nv4) {
            vq[k] = ld4(q + 4 * c);
            if (l2) {
                const float dx = vp[k].x - vq[k].x, dy = vp[k].y - vq[k].y, dz = vp[k].z - vq[k].z,
                            dw = vp[k].w - vq[k].w;
                part -= dx * dx + dy * dy + dz * dz + dw * dw;  // warp.cc:25-28
            } else {
                part += vp[k].x * vq[k].x + vp[k].y * vq[k].y + vp[k].z * vq[k].z + vp[k].w * vq[k].w;  // :21-23
            }
        }
    }
    return warp_sum(part);
}

template <int NV>
__global__ void __launch_bounds__(256) warp_accumulate_kernel(SgdArgs a) {
    const int lane = threadIdx.x & 31;
    const int64_t w0 = (int64_t)blockIdx.x * (blockDim.x >> 5) + warp_id_uniform();
    const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int nv4 = a.ld >> 2;
    double loss = 0.0;
    unsigned long long updates = 0;
    for (int64_t it = a.it_begin + w0; it < a.it_end; it += nw) {
        const int64_t row = uni((long long)row_of(a.indptr, a.row_begin, a.row_end, it));
        const int64_t beg = uni((long long)(row == 0 ? 0 : __ldg(a.indptr + row - 1)));
        const int64_t end = uni((long long)__ldg(a.indptr + row));
        const int32_t* rk = a.keys + (beg - a.shift);
        const int64_t n_seen = end - beg;
        const int pos = uni(__ldg(a.keys + (it - a.shift)));
        const float* pu = a.P + row * a.ld;
        const float* qi = a.Q + (int64_t)pos * a.ld;
        float4 vp[NV], vi[NV], vj[NV];
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            const int c = lane + 32 * k;
            vp[k] = c < nv4 ? ld4(pu + 4 * c) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        const float ui = uni(warp_score<NV>(vp, qi, nv4, lane, a.score_l2, vi));  // warp.cc:133
        float uj = 0.f;
        int neg = 0;
        int trial = 1;
        uint32_t t = 0;
        while (trial <= a.max_trials) {  // warp.cc:137-148
            neg = draw_range(a.seed, a.epoch, (uint64_t)it, t++, (uint32_t)a.num_items);
            if (uni(seen_sorted(rk, n_seen, neg))) {  // :140-141, not counted as a trial
                if (t > (uint32_t)(64 * a.max_trials + 4096)) {
                    trial = a.max_trials + 1;
                    break;
                }
                continue;
            }
            trial += 1;  // :142
            uj = uni(warp_score<NV>(vp, a.Q + (int64_t)neg * a.ld, nv4, lane, a.score_l2, vj));
            if ((ui - uj) < a.threshold) break;  // :145-146
            trial += 1;  // :147
        }
        const bool discard = trial >= a.max_trials;  // :149-150
        if (lane == 0) {
            if (a.trace_trials) a.trace_trials[it - a.shift] = discard ? 0 : trial;
            if (a.trace_negs) a.trace_negs[it - a.shift] = discard ? -1 : neg;
        }
        if (discard) continue;
        int64_t ratio = ((int64_t)a.num_items - n_seen - 1) / trial;  // :152
        if (ratio < 1) ratio = 1;
        const float Phi = logf((float)(int)ratio);
        float* gp = a.gP + row * a.ld;
        float* gi = a.gQ + (int64_t)pos * a.ld;
        float* gj = a.gQ + (int64_t)neg * a.ld;
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            const int c = lane + 32 * k;
            if (c < nv4) {
                float4 du, di, dj;
                if (!a.score_l2) {  // dot_deriv warp.cc:30-40
                    du = make_float4(Phi * (vi[k].x - vj[k].x), Phi * (vi[k].y - vj[k].y), Phi * (vi[k].z - vj[k].z),
                                     Phi * (vi[k].w - vj[k].w));
                    di = make_float4(Phi * vp[k].x, Phi * vp[k].y, Phi * vp[k].z, Phi * vp[k].w);
                    dj = make_float4(-di.x, -di.y, -di.z, -di.w);
                } else {  // l2_deriv warp.cc:42-52
                    du = make_float4(Phi * 2 * (vi[k].x - vj[k].x), Phi * 2 * (vi[k].y - vj[k].y),
                                     Phi * 2 * (vi[k].z - vj[k].z), Phi * 2 * (vi[k].w - vj[k].w));
                    di = make_float4(Phi * (vp[k].x - vi[k].x), Phi * (vp[k].y - vi[k].y), Phi * (vp[k].z - vi[k].z),
                                     Phi * (vp[k].w - vi[k].w));
                    dj = make_float4(-Phi * (vp[k].x - vj[k].x), -Phi * (vp[k].y - vj[k].y),
                                     -Phi * (vp[k].z - vj[k].z), -Phi * (vp[k].w - vj[k].w));
                }
                // grad += deriv - reg * param  (warp.cc:156-158)
                red4(gp + 4 * c, make_float4(du.x - a.reg_u * vp[k].x, du.y - a.reg_u * vp[k].y,
                                             du.z - a.reg_u * vp[k].z, du.w - a.reg_u * vp[k].w));
                red4(gi + 4 * c, make_float4(di.x - a.reg_i * vi[k].x, di.y - a.reg_i * vi[k].y,
                                             di.z - a.reg_i * vi[k].z, di.w - a.reg_i * vi[k].w));
                red4(gj + 4 * c, make_float4(dj.x - a.reg_j * vj[k].x, dj.y - a.reg_j * vj[k].y,
                                             dj.z - a.reg_j * vj[k].z, dj.w - a.reg_j * vj[k].w));
            }
        }
        if (lane == 0 && a.pcn) {  // warp.cc:159-165
            atomicAdd(a.cP + row, 1);
            atomicAdd(a.cQ + pos, 1);
            atomicAdd(a.cQ + neg, 1);
        }
        loss += (double)(uj - ui + a.threshold);  // :166
        updates += 1;
    }
    if (lane == 0 && updates) {
        atomicAdd(a.stat_loss, loss);
        atomicAdd(a.stat_updates, updates);
    }
}

// ---------------------------------------------------------------------------------------
// SGDAlgorithm::update_parameters (algo.cc:382-465): element-wise Adam / Adagrad step.
// beta2 := beta1 (algo.cc:396); the gradient buffer ends up holding the step and is NOT
// cleared (no setZero in the reference).  One thread per element; `cols` = row pitch.
// ---------------------------------------------------------------------------------------
__global__ void sgd_apply_kernel(int optimizer, float* __restrict__ theta, float* __restrict__ grad,
                                 float* __restrict__ mom, float* __restrict__ vel, const int32_t* __restrict__ cnt,
                                 int64_t rows, int cols, float two_reg, float lr, float b1, float omb1, float b2,
                                 float omb2, float bc1, float bc2, int pcn) {
    const int64_t n = rows * cols;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
        float g = grad[e];
        const float th = theta[e];
        if (pcn) {
            const int c = cnt[e / cols];
            if (c) g /= (float)c;  // algo.cc:399-401
        }
        g -= th * two_reg;  // :403
        if (optimizer == 2) {  // update_adam :365-375
            const float m = b1 * mom[e] + omb1 * g;
            const float v = b2 * vel[e] + omb2 * (g * g);
            mom[e] = m;
            vel[e] = v;
            g = (m / bc1) / (sqrtf(v / bc2) + 1e-10f);
        } else {  // update_adagrad :377-380
            const float v = vel[e] + g * g;
            vel[e] = v;
            g = g / (sqrtf(v) + 1e-10f);
        }
        grad[e] = g;
        theta[e] = th + lr * g;  // :405
    }
}

// CWARP::update_parameters tail (warp.cc:194-200): row /= max(1, ||row||).  One warp per row.
__global__ void warp_project_kernel(float* __restrict__ M, int64_t rows, int ld) {
    const int lane = threadIdx.x & 31;
    const int64_t w0 = (int64_t)blockIdx.x * (blockDim.x >> 5) + warp_id_uniform();
    const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r = w0; r < rows; r += nw) {
        float* m = M + r * ld;
        float s = 0.f;
        for (int c = lane; c < ld; c += 32) s += m[c] * m[c];
        s = sqrtf(warp_sum(s));
        if (s > 1.0f)
            for (int c = lane; c < ld; c += 32) m[c] /= s;
    }
}

// probe losses: BPR mean log(1+exp(-x_uij)) (bpr.cc:227-244); WARP fraction violating (warp.cc:205-226)
__global__ void probe_loss_kernel(int kind, const float* __restrict__ P, const float* __restrict__ Q,
                                  const float* __restrict__ Qb, int D, int ld, int use_bias, int l2,
                                  double threshold, const int32_t* __restrict__ us, const int32_t* __restrict__ ps,
                                  const int32_t* __restrict__ ns, int n, double* out) {
    const int lane = threadIdx.x & 31;
    const int w0 = blockIdx.x * (blockDim.x >> 5) + warp_id_uniform();
    const int nw = (gridDim.x * blockDim.x) >> 5;
    double acc = 0.0;
    for (int i = w0; i < n; i += nw) {
        const float* p = P + (int64_t)us[i] * ld;
        const float* qi = Q + (int64_t)ps[i] * ld;
        const float* qj = Q + (int64_t)ns[i] * ld;
        float sa = 0.f, sb = 0.f;
        for (int c = lane; c < D; c += 32) {
            if (l2) {
                const float da = p[c] - qi[c], db = p[c] - qj[c];
                sa -= da * da;
                sb -= db * db;
            } else {
                sa += p[c] * qi[c];
                sb += p[c] * qj[c];
            }
        }
        sa = warp_sum(sa);
        sb = warp_sum(sb);
        if (kind == BFL_SGD_BPR) {
            if (use_bias) {
                sa += Qb[ps[i]];
                sb += Qb[ns[i]];
            }
            acc += log(1.0 + exp(-((double)sa - (double)sb)));
        } else {
            acc += (((double)sa - (double)sb) < threshold) ? 1.0 : 0.0;
        }
    }
    if (lane == 0 && acc != 0.0) atomicAdd(out, acc);
}

}  // namespace

struct bfl_sgd {
    int kind = BFL_SGD_BPR;
    bool opt_set = false;
    int d = 0, vdim = 0;
    int optimizer = 0;
    bool use_bias = true, update_i = true, update_j = true, verify_neg = true, uniform = true, pcn = false;
    bool compute_loss = true, score_l2 = false;
    int num_neg = 1, max_trials = 500, num_iters = 1;
    uint32_t seed = 0;
    float reg_u = 0, reg_i = 0, reg_j = 0, reg_b = 0, threshold = 1.f;
    double lr0 = 0.05, min_lr = 1e-4, beta1 = 0.9;

    float *hostP = nullptr, *hostQ = nullptr, *hostQb = nullptr;
    DevBuf<float> ownP, ownQ, ownQb;
    float *dP = nullptr, *dQ = nullptr, *dQb = nullptr;
    int64_t P_rows = 0, Q_rows = 0;
    bool factors_ready = false;
    DevBuf<float> gP, gQ, gQb, mP, mQ, mQb, vP, vQ, vQb;
    DevBuf<int32_t> cP, cQ;
    DevBuf<int64_t> cum;
    bool cum_set = false;

    DevBuf<int64_t> own_indptr;
    const int64_t* d_indptr = nullptr;
    DevBuf<int32_t> stage_keys;
    const int32_t* d_keys = nullptr;
    int64_t csr_rows = 0, csr_nnz = 0;
    DevBuf<int32_t> tri_u, tri_p, tri_n;
    DevBuf<int32_t> probe;
    DevBuf<double> d_stat;  // [0] warp loss sum, [1] probe loss
    DevBuf<unsigned long long> d_upd;
    int32_t* trace_trials = nullptr;
    int32_t* trace_negs = nullptr;

    int iters = 0, epoch = 0;
    double processed = 0.0, total = 1.0, cur_lr = 0.0;
    cudaStream_t stream = nullptr;
    int num_sms = 148;
};

namespace {

int sgd_apply_options(bfl_sgd* h, const JsonOpt& j) {
    h->d = j.integer("d", h->kind == BFL_SGD_WARP ? 64 : 20);
    if (h->d <= 0 || h->d > 512) BFL_FAIL(BFL_ERR_OPTION, "d must be in [1, 512]");
    h->vdim = (h->d + 3) / 4 * 4;
    std::string optimizer = j.string("optimizer", h->kind == BFL_SGD_WARP ? "adagrad" : "sgd");
    if (optimizer == "sgd") h->optimizer = 0;
    else if (optimizer == "adagrad") h->optimizer = 1;
    else if (optimizer == "adam") h->optimizer = 2;
    else BFL_FAIL(BFL_ERR_OPTION, "optimizer must be one of sgd, adagrad, adam");
    if (h->kind == BFL_SGD_WARP && h->optimizer == 0)
        BFL_FAIL(BFL_ERR_OPTION, "WARP accumulates gradients only (warp.cc:156-158): optimizer must be adagrad or adam");
    h->use_bias = h->kind == BFL_SGD_WARP ? false : j.flag("use_bias", true);
    h->update_i = j.flag("update_i", true);
    h->update_j = j.flag("update_j", true);
    h->verify_neg = j.flag("verify_neg", true);
    h->uniform = j.number("sampling_power", 0.0) == 0.0;  // bpr.cc:91
    h->pcn = j.flag("per_coordinate_normalize", false);
    h->compute_loss = j.flag("compute_loss_on_training", true);
    std::string sf = j.string("score_func", "dot");
    h->score_l2 = (sf == "l2" || sf == "L2");
    h->num_neg = j.integer("num_negative_samples", 1);
    if (h->num_neg < 1) h->num_neg = 1;
    h->max_trials = j.integer("max_trials", 500);
    h->num_iters = j.integer("num_iters", 1);
    h->seed = (uint32_t)j.integer("random_seed", 0);
    h->reg_u = (float)j.number("reg_u", 0.0);
    h->reg_i = (float)j.number("reg_i", 0.0);
    h->reg_j = (float)j.number("reg_j", 0.0);
    h->reg_b = (float)j.number("reg_b", 0.0);
    h->threshold = (float)j.number("threshold", 1.0);
    h->lr0 = j.number("lr", 0.05);
    h->min_lr = j.number("min_lr", 0.0001);
    h->beta1 = j.number("beta1", 0.9);
    if (BFL_OK != require_device()) return BFL_ERR_CUDA;
    int dev = 0;
    BFL_CUDA(cudaGetDevice(&dev));
    BFL_CUDA(cudaDeviceGetAttribute(&h->num_sms, cudaDevAttrMultiProcessorCount, dev));
    if (!h->stream) BFL_CUDA(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    if (BFL_OK != h->d_stat.reserve(2)) return BFL_ERR_CUDA;
    if (BFL_OK != h->d_upd.reserve(1)) return BFL_ERR_CUDA;
    BFL_CUDA(cudaMemsetAsync(h->d_stat.p, 0, 2 * sizeof(double), h->stream));
    BFL_CUDA(cudaMemsetAsync(h->d_upd.p, 0, sizeof(unsigned long long), h->stream));
    h->cur_lr = h->lr0;
    h->opt_set = true;
    return BFL_OK;
}

int alloc_state(bfl_sgd* h, int64_t num_total_samples) {
    const size_t np = (size_t)h->P_rows * h->vdim, nq = (size_t)h->Q_rows * h->vdim, nb = (size_t)h->Q_rows;
    cudaStream_t st = h->stream;
    if (h->optimizer != 0) {  // initialize_adam_optimizer algo.cc:221-254
        DevBuf<float>* bufs[] = {&h->gP, &h->gQ, &h->gQb, &h->mP, &h->mQ, &h->mQb, &h->vP, &h->vQ, &h->vQb};
        const size_t sizes[] = {np, nq, nb, np, nq, nb, np, nq, nb};
        for (int i = 0; i < 9; ++i) {
            if (h->optimizer == 1 && i >= 3 && i < 6) continue;  // adagrad needs no momentum
            if (BFL_OK != bufs[i]->reserve(sizes[i])) return BFL_ERR_CUDA;
            BFL_CUDA(cudaMemsetAsync(bufs[i]->p, 0, sizes[i] * sizeof(float), st));
        }
    }
    if (BFL_OK != h->cP.reserve((size_t)h->P_rows)) return BFL_ERR_CUDA;
    if (BFL_OK != h->cQ.reserve((size_t)h->Q_rows)) return BFL_ERR_CUDA;
    BFL_CUDA(cudaMemsetAsync(h->cP.p, 0, sizeof(int32_t) * h->P_rows, st));
    BFL_CUDA(cudaMemsetAsync(h->cQ.p, 0, sizeof(int32_t) * h->Q_rows, st));
    h->iters = 0;
    h->epoch = 0;
    h->processed = 0.0;
    h->total = (double)num_total_samples * (double)h->num_iters;  // algo.cc:174-175
    if (h->total <= 0) h->total = 1.0;
    h->cur_lr = h->lr0;
    BFL_CUDA(cudaStreamSynchronize(st));
    h->factors_ready = true;
    return BFL_OK;
}

void fill_args(bfl_sgd* h, SgdArgs& a, const int32_t* keys, int64_t shift, int64_t row_begin, int64_t row_end,
               int64_t it_begin, int64_t it_end) {
    a.P = h->dP; a.Q = h->dQ; a.Qb = h->dQb;
    a.gP = h->gP.p; a.gQ = h->gQ.p; a.gQb = h->gQb.p;
    a.cP = h->cP.p; a.cQ = h->cQ.p;
    a.indptr = h->d_indptr; a.keys = keys;
    a.cum = h->cum_set ? h->cum.p : nullptr;
    a.trace_trials = h->trace_trials; a.trace_negs = h->trace_negs;
    a.stat_loss = h->d_stat.p; a.stat_updates = h->d_upd.p;
    a.shift = shift; a.row_begin = row_begin; a.row_end = row_end; a.it_begin = it_begin; a.it_end = it_end;
    a.num_items = (int32_t)h->Q_rows; a.D = h->d; a.ld = h->vdim;
    a.optimizer = h->optimizer; a.use_bias = h->use_bias; a.update_i = h->update_i; a.update_j = h->update_j;
    a.num_neg = h->num_neg; a.verify_neg = h->verify_neg; a.uniform = h->uniform || !h->cum_set; a.pcn = h->pcn;
    a.max_trials = h->max_trials; a.score_l2 = h->score_l2; a.seed = h->seed; a.epoch = (uint32_t)h->epoch;
    a.reg_u = h->reg_u; a.reg_i = h->reg_i; a.reg_j = h->reg_j; a.reg_b = h->reg_b;
    a.lr = (float)h->cur_lr; a.threshold = h->threshold;
}

int launch_bpr_apply(bfl_sgd* h, const SgdArgs& a, const int32_t* u, const int32_t* p, const int32_t* n, int64_t cnt,
                     cudaStream_t st) {
    if (cnt <= 0) return BFL_OK;
    // >= 16 consecutive triples per warp, at most 32 warps per SM (staleness grows with the number of triples in flight)
    const int grid = (int)std::min<int64_t>((cnt + 127) / 128, (int64_t)h->num_sms * 4);
    const int nv = (h->vdim / 4 + 31) / 32;
    if (nv <= 1) bpr_apply_kernel<1><<<grid, 256, 0, st>>>(a, u, p, n, cnt);
    else if (nv <= 2) bpr_apply_kernel<2><<<grid, 256, 0, st>>>(a, u, p, n, cnt);
    else bpr_apply_kernel<4><<<grid, 256, 0, st>>>(a, u, p, n, cnt);
    BFL_LAUNCHED();
    return BFL_OK;
}

// process rows [row_begin,row_end) whose positives are keys[it - shift], it in [it_begin, it_end)
int run_jobs(bfl_sgd* h, const int32_t* keys, int64_t shift, int64_t row_begin, int64_t row_end, int64_t it_begin,
             int64_t it_end, cudaStream_t st) {
    const int64_t npos = it_end - it_begin;
    // job.alpha = lr_ at job creation (algo.cc:351,359); linear decay by processed fraction (:284-287)
    double lr = h->lr0 - (h->lr0 - h->min_lr) * (h->processed / h->total);
    h->cur_lr = lr > h->min_lr ? lr : h->min_lr;
    if (npos <= 0) return BFL_OK;
    SgdArgs a;
    fill_args(h, a, keys, shift, row_begin, row_end, it_begin, it_end);
    if (h->kind == BFL_SGD_WARP) {
        const int grid = (int)std::min<int64_t>((npos + 7) / 8, (int64_t)h->num_sms * 16);
        const int nv = (h->vdim / 4 + 31) / 32;
        if (nv <= 1) warp_accumulate_kernel<1><<<grid, 256, 0, st>>>(a);
        else if (nv <= 2) warp_accumulate_kernel<2><<<grid, 256, 0, st>>>(a);
        else warp_accumulate_kernel<4><<<grid, 256, 0, st>>>(a);
        BFL_LAUNCHED();
    } else {
        // sample + apply in slabs so the triple buffers stay bounded (<= 64M samples)
        const int64_t slab_pos = std::max<int64_t>(1, (int64_t)(1 << 26) / h->num_neg);
        for (int64_t b = it_begin; b < it_end; b += slab_pos) {
            const int64_t e = std::min(it_end, b + slab_pos);
            const int64_t cnt = (e - b) * h->num_neg;
            if (BFL_OK != h->tri_u.reserve((size_t)cnt)) return BFL_ERR_CUDA;
            if (BFL_OK != h->tri_p.reserve((size_t)cnt)) return BFL_ERR_CUDA;
            if (BFL_OK != h->tri_n.reserve((size_t)cnt)) return BFL_ERR_CUDA;
            SgdArgs s = a;
            s.it_begin = b;
            s.it_end = e;
            const int grid = (int)std::min<int64_t>((cnt + 255) / 256, (int64_t)h->num_sms * 32);
            bpr_sample_kernel<<<grid, 256, 0, st>>>(s, h->tri_u.p, h->tri_p.p, h->tri_n.p);
            BFL_LAUNCHED();
            int rc = launch_bpr_apply(h, s, h->tri_u.p, h->tri_p.p, h->tri_n.p, cnt, st);
            if (rc != BFL_OK) return rc;
        }
    }
    h->processed += (double)npos;
    return BFL_OK;
}

int apply_optimizer(bfl_sgd* h, cudaStream_t st) {
    if (h->optimizer != 0) {
        const double beta2 = h->beta1;  // algo.cc:396
        const float b1 = (float)h->beta1, omb1 = (float)(1.0 - h->beta1);
        const float b2 = (float)beta2, omb2 = (float)(1.0 - beta2);
        const float bc1 = (float)(1.0 - pow(h->beta1, h->iters + 1));
        const float bc2 = (float)(1.0 - pow(beta2, h->iters + 1));
        struct Item { float* th; float* g; float* m; float* v; const int32_t* c; int64_t rows; int cols; double reg; };
        Item items[3] = {{h->dP, h->gP.p, h->mP.p, h->vP.p, h->cP.p, h->P_rows, h->vdim, h->reg_u},
                         {h->dQ, h->gQ.p, h->mQ.p, h->vQ.p, h->cQ.p, h->Q_rows, h->vdim, h->reg_i},
                         {h->dQb, h->gQb.p, h->mQb.p, h->vQb.p, h->cQ.p, h->Q_rows, 1, h->reg_b}};
        const int nitems = (h->use_bias && h->kind == BFL_SGD_BPR) ? 3 : 2;
        for (int i = 0; i < nitems; ++i) {
            const Item& it = items[i];
            const int64_t n = it.rows * it.cols;
            const int grid = (int)std::min<int64_t>((n + 255) / 256, (int64_t)h->num_sms * 32);
            sgd_apply_kernel<<<grid, 256, 0, st>>>(h->optimizer, it.th, it.g, it.m, it.v, it.c, it.rows, it.cols,
                                                   (float)(2 * it.reg), (float)h->lr0, b1, omb1, b2, omb2, bc1, bc2,
                                                   h->pcn ? 1 : 0);
            BFL_LAUNCHED();
        }
        if (h->pcn) {  // algo.cc:424-427
            BFL_CUDA(cudaMemsetAsync(h->cP.p, 0, sizeof(int32_t) * h->P_rows, st));
            BFL_CUDA(cudaMemsetAsync(h->cQ.p, 0, sizeof(int32_t) * h->Q_rows, st));
        }
    }
    if (h->kind == BFL_SGD_WARP) {  // warp.cc:192-201
        const int gq = (int)std::min<int64_t>((h->Q_rows + 7) / 8, (int64_t)h->num_sms * 16);
        warp_project_kernel<<<gq, 256, 0, st>>>(h->dQ, h->Q_rows, h->vdim);
        BFL_LAUNCHED();
        const int gp = (int)std::min<int64_t>((h->P_rows + 7) / 8, (int64_t)h->num_sms * 16);
        warp_project_kernel<<<gp, 256, 0, st>>>(h->dP, h->P_rows, h->vdim);
        BFL_LAUNCHED();
    }
    h->iters += 1;  // algo.cc:464
    h->epoch += 1;
    return BFL_OK;
}

int sync_to_host(bfl_sgd* h) {
    if (!h->hostP) return BFL_OK;
    BFL_CUDA(cudaMemcpyAsync(h->hostP, h->dP, sizeof(float) * (size_t)h->P_rows * h->vdim, cudaMemcpyDeviceToHost, h->stream));
    BFL_CUDA(cudaMemcpyAsync(h->hostQ, h->dQ, sizeof(float) * (size_t)h->Q_rows * h->vdim, cudaMemcpyDeviceToHost, h->stream));
    BFL_CUDA(cudaMemcpyAsync(h->hostQb, h->dQb, sizeof(float) * (size_t)h->Q_rows, cudaMemcpyDeviceToHost, h->stream));
    BFL_CUDA(cudaStreamSynchronize(h->stream));
    return BFL_OK;
}

}  // namespace

extern "C" {

bfl_sgd_t* bfl_sgd_create(int kind) {
    if (kind != BFL_SGD_BPR && kind != BFL_SGD_WARP) return nullptr;
    bfl_sgd* h = new (std::nothrow) bfl_sgd();
    if (h) h->kind = kind;
    return h;
}

void bfl_sgd_destroy(bfl_sgd_t* h) {
    if (!h) return;
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
}

int bfl_sgd_init(bfl_sgd_t* h, const char* opt_path) {
    if (!h || !opt_path) BFL_FAIL(BFL_ERR_ARG, "null argument");
    JsonOpt j;
    std::string err;
    if (!j.load(opt_path, &err)) BFL_FAIL(BFL_ERR_OPTION, err);
    return sgd_apply_options(h, j);
}

int bfl_sgd_init_json(bfl_sgd_t* h, const char* json_text) {
    if (!h || !json_text) BFL_FAIL(BFL_ERR_ARG, "null argument");
    JsonOpt j;
    std::string err;
    if (!j.parse(json_text, &err)) BFL_FAIL(BFL_ERR_OPTION, "Failed to parse: " + err);
    return sgd_apply_options(h, j);
}

int bfl_sgd_get_vdim(bfl_sgd_t* h) { return h ? h->vdim : 0; }

int bfl_sgd_initialize_model(bfl_sgd_t* h, float* P, int32_t P_rows, float* Q, int32_t Q_rows, float* Qb,
                             int64_t num_total_samples) {
    if (!h || !h->opt_set) BFL_FAIL(BFL_ERR_STATE, "init() must succeed before initialize_model()");
    if (!P || !Q || !Qb || P_rows <= 0 || Q_rows <= 0) BFL_FAIL(BFL_ERR_ARG, "bad factor arguments");
    h->hostP = P; h->hostQ = Q; h->hostQb = Qb;
    h->P_rows = P_rows; h->Q_rows = Q_rows;
    if (BFL_OK != h->ownP.reserve((size_t)P_rows * h->vdim)) return BFL_ERR_CUDA;
    if (BFL_OK != h->ownQ.reserve((size_t)Q_rows * h->vdim)) return BFL_ERR_CUDA;
    if (BFL_OK != h->ownQb.reserve((size_t)Q_rows)) return BFL_ERR_CUDA;
    h->dP = h->ownP.p; h->dQ = h->ownQ.p; h->dQb = h->ownQb.p;
    BFL_CUDA(cudaMemcpyAsync(h->dP, P, sizeof(float) * (size_t)P_rows * h->vdim, cudaMemcpyHostToDevice, h->stream));
    BFL_CUDA(cudaMemcpyAsync(h->dQ, Q, sizeof(float) * (size_t)Q_rows * h->vdim, cudaMemcpyHostToDevice, h->stream));
    BFL_CUDA(cudaMemcpyAsync(h->dQb, Qb, sizeof(float) * (size_t)Q_rows, cudaMemcpyHostToDevice, h->stream));
    return alloc_state(h, num_total_samples);
}

int bfl_sgd_bind_factors_device(bfl_sgd_t* h, float* dP, int64_t P_rows, float* dQ, int64_t Q_rows, float* dQb,
                                int64_t num_total_samples) {
    if (!h || !h->opt_set) BFL_FAIL(BFL_ERR_STATE, "init() must succeed before binding factors");
    if (!dP || !dQ || !dQb || P_rows <= 0 || Q_rows <= 0) BFL_FAIL(BFL_ERR_ARG, "bad factor arguments");
    if (((uintptr_t)dP | (uintptr_t)dQ) & 15) BFL_FAIL(BFL_ERR_ARG, "device factor pointers must be 16-byte aligned");
    h->hostP = h->hostQ = h->hostQb = nullptr;
    h->ownP.release(); h->ownQ.release(); h->ownQb.release();
    h->dP = dP; h->dQ = dQ; h->dQb = dQb;
    h->P_rows = P_rows; h->Q_rows = Q_rows;
    return alloc_state(h, num_total_samples);
}

int bfl_sgd_set_cumulative_table(bfl_sgd_t* h, const int64_t* cum_table, int32_t size) {
    if (!h || !h->opt_set) BFL_FAIL(BFL_ERR_STATE, "init() must precede set_cumulative_table()");
    if (!cum_table || size <= 0) BFL_FAIL(BFL_ERR_ARG, "bad cumulative table");
    if (BFL_OK != h->cum.reserve((size_t)size)) return BFL_ERR_CUDA;
    BFL_CUDA(cudaMemcpyAsync(h->cum.p, cum_table, sizeof(int64_t) * size, cudaMemcpyHostToDevice, h->stream));
    BFL_CUDA(cudaStreamSynchronize(h->stream));
    // an all-zero table (sampling_power == 0, bpr.py:101-111) means uniform sampling
    h->cum_set = cum_table[size - 1] > 0;
    return BFL_OK;
}

int bfl_sgd_set_placeholder(bfl_sgd_t* h, const int64_t* indptr, size_t batch_size) {
    if (!h || !h->factors_ready) BFL_FAIL(BFL_ERR_STATE, "initialize_model() must precede set_placeholder()");
    if (!indptr) BFL_FAIL(BFL_ERR_ARG, "null indptr");
    if (BFL_OK != h->own_indptr.reserve((size_t)h->P_rows)) return BFL_ERR_CUDA;
    BFL_CUDA(cudaMemcpyAsync(h->own_indptr.p, indptr, sizeof(int64_t) * h->P_rows, cudaMemcpyHostToDevice, h->stream));
    h->d_indptr = h->own_indptr.p;
    if (batch_size && BFL_OK != h->stage_keys.reserve(batch_size)) return BFL_ERR_CUDA;
    BFL_CUDA(cudaStreamSynchronize(h->stream));
    return BFL_OK;
}

int bfl_sgd_bind_csr_device(bfl_sgd_t* h, const int64_t* d_indptr, const int32_t* d_keys, int64_t rows, int64_t nnz) {
    if (!h || !h->opt_set) BFL_FAIL(BFL_ERR_STATE, "init() must precede bind_csr");
    if (!d_indptr || (nnz > 0 && !d_keys) || rows <= 0) BFL_FAIL(BFL_ERR_ARG, "bad CSR arguments");
    h->d_indptr = d_indptr;
    h->d_keys = d_keys;
    h->csr_rows = rows;
    h->csr_nnz = nnz;
    return BFL_OK;
}

int bfl_sgd_launch_workers(bfl_sgd_t* h) {
    if (!h || !h->factors_ready) BFL_FAIL(BFL_ERR_STATE, "initialize_model() must precede launch_workers()");
    return BFL_OK;
}

int bfl_sgd_wait_until_done(bfl_sgd_t* h) {
    if (!h) BFL_FAIL(BFL_ERR_ARG, "null handle");
    if (h->stream) BFL_CUDA(cudaStreamSynchronize(h->stream));
    return BFL_OK;
}

int bfl_sgd_join(bfl_sgd_t* h, double* out) {
    if (out) *out = 0.0;
    if (!h) BFL_FAIL(BFL_ERR_ARG, "null handle");
    if (h->stream) BFL_CUDA(cudaStreamSynchronize(h->stream));
    return sync_to_host(h);
}

int bfl_sgd_add_jobs(bfl_sgd_t* h, int32_t start_x, int32_t next_x, const int64_t* indptr, const int32_t* keys) {
    if (!h || !h->factors_ready) BFL_FAIL(BFL_ERR_STATE, "initialize_model() must precede add_jobs()");
    if (next_x - start_x == 0) return BFL_OK;  // algo.cc:314-317
    if (start_x < 0 || next_x > h->P_rows || next_x < start_x || !indptr || !keys) BFL_FAIL(BFL_ERR_ARG, "bad chunk arguments");
    if (h->d_indptr != h->own_indptr.p || !h->own_indptr.p) {
        if (BFL_OK != h->own_indptr.reserve((size_t)h->P_rows)) return BFL_ERR_CUDA;
        BFL_CUDA(cudaMemcpyAsync(h->own_indptr.p, indptr, sizeof(int64_t) * h->P_rows, cudaMemcpyHostToDevice, h->stream));
        h->d_indptr = h->own_indptr.p;
    }
    const int64_t beg = start_x == 0 ? 0 : indptr[start_x - 1];
    const int64_t end = indptr[next_x - 1];
    const int64_t n = end - beg;
    if (n > 0) {
        if (BFL_OK != h->stage_keys.reserve((size_t)n)) return BFL_ERR_CUDA;
        BFL_CUDA(cudaMemcpyAsync(h->stage_keys.p, keys, sizeof(int32_t) * n, cudaMemcpyHostToDevice, h->stream));
    }
    int rc = run_jobs(h, h->stage_keys.p, beg, start_x, next_x, beg, end, h->stream);
    if (rc != BFL_OK) return rc;
    // the staging buffer is reused by the next chunk: drain before returning
    BFL_CUDA(cudaStreamSynchronize(h->stream));
    return BFL_OK;
}

int bfl_sgd_add_jobs_device(bfl_sgd_t* h, int64_t row_begin, int64_t row_end, void* stream) {
    if (!h || !h->factors_ready) BFL_FAIL(BFL_ERR_STATE, "factors not bound");
    if (!h->d_indptr || (!h->d_keys && h->csr_nnz > 0)) BFL_FAIL(BFL_ERR_STATE, "no device CSR bound");
    if (row_begin < 0 || row_end > h->csr_rows || row_end < row_begin) BFL_FAIL(BFL_ERR_ARG, "bad row range");
    if (row_end == row_begin) return BFL_OK;
    int64_t ends[2] = {0, 0};
    if (row_begin > 0)
        BFL_CUDA(cudaMemcpyAsync(&ends[0], h->d_indptr + row_begin - 1, sizeof(int64_t), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    BFL_CUDA(cudaMemcpyAsync(&ends[1], h->d_indptr + row_end - 1, sizeof(int64_t), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    BFL_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    return run_jobs(h, h->d_keys, 0, row_begin, row_end, ends[0], ends[1], (cudaStream_t)stream);
}

int bfl_sgd_update_parameters(bfl_sgd_t* h) {
    if (!h || !h->factors_ready) BFL_FAIL(BFL_ERR_STATE, "initialize_model() must precede update_parameters()");
    int rc = apply_optimizer(h, h->stream);
    if (rc != BFL_OK) return rc;
    return sync_to_host(h);  // cuda/_bpr.pyx:60-61
}

int bfl_sgd_update_parameters_device(bfl_sgd_t* h, void* stream) {
    if (!h || !h->factors_ready) BFL_FAIL(BFL_ERR_STATE, "factors not bound");
    return apply_optimizer(h, (cudaStream_t)stream);
}

int bfl_sgd_synchronize(bfl_sgd_t* h, int device_to_host) {
    if (!h || !h->factors_ready) BFL_FAIL(BFL_ERR_STATE, "initialize_model() must precede synchronize()");
    if (device_to_host) return sync_to_host(h);
    if (!h->hostP) return BFL_OK;
    BFL_CUDA(cudaMemcpyAsync(h->dP, h->hostP, sizeof(float) * (size_t)h->P_rows * h->vdim, cudaMemcpyHostToDevice, h->stream));
    BFL_CUDA(cudaMemcpyAsync(h->dQ, h->hostQ, sizeof(float) * (size_t)h->Q_rows * h->vdim, cudaMemcpyHostToDevice, h->stream));
    BFL_CUDA(cudaMemcpyAsync(h->dQb, h->hostQb, sizeof(float) * (size_t)h->Q_rows, cudaMemcpyHostToDevice, h->stream));
    BFL_CUDA(cudaStreamSynchronize(h->stream));
    return BFL_OK;
}

int bfl_sgd_compute_loss(bfl_sgd_t* h, int32_t n, const int32_t* users, const int32_t* positives,
                         const int32_t* negatives, double* out_loss) {
    if (out_loss) *out_loss = 0.0;
    if (!h || !h->factors_ready) BFL_FAIL(BFL_ERR_STATE, "initialize_model() must precede compute_loss()");
    if (n <= 0) return BFL_OK;
    if (!users || !positives || !negatives) BFL_FAIL(BFL_ERR_ARG, "null probe arrays");
    if (BFL_OK != h->probe.reserve((size_t)3 * n)) return BFL_ERR_CUDA;
    BFL_CUDA(cudaMemcpyAsync(h->probe.p, users, sizeof(int32_t) * n, cudaMemcpyHostToDevice, h->stream));
    BFL_CUDA(cudaMemcpyAsync(h->probe.p + n, positives, sizeof(int32_t) * n, cudaMemcpyHostToDevice, h->stream));
    BFL_CUDA(cudaMemcpyAsync(h->probe.p + 2 * n, negatives, sizeof(int32_t) * n, cudaMemcpyHostToDevice, h->stream));
    BFL_CUDA(cudaMemsetAsync(h->d_stat.p + 1, 0, sizeof(double), h->stream));
    const int grid = std::min((n + 7) / 8, h->num_sms * 4);
    probe_loss_kernel<<<grid, 256, 0, h->stream>>>(h->kind, h->dP, h->dQ, h->dQb, h->d, h->vdim, h->use_bias,
                                                   h->score_l2, (double)h->threshold, h->probe.p, h->probe.p + n,
                                                   h->probe.p + 2 * n, n, h->d_stat.p + 1);
    BFL_LAUNCHED();
    double v = 0.0;
    BFL_CUDA(cudaMemcpyAsync(&v, h->d_stat.p + 1, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    BFL_CUDA(cudaStreamSynchronize(h->stream));
    if (out_loss) *out_loss = v / (double)n;
    return BFL_OK;
}

int bfl_sgd_apply_triples_device(bfl_sgd_t* h, const int32_t* d_users, const int32_t* d_pos, const int32_t* d_neg,
                                 int64_t n, float lr, void* stream) {
    if (!h || !h->factors_ready) BFL_FAIL(BFL_ERR_STATE, "factors not bound");
    if (h->kind != BFL_SGD_BPR) BFL_FAIL(BFL_ERR_STATE, "explicit triples are a BPR hook");
    SgdArgs a;
    fill_args(h, a, nullptr, 0, 0, 0, 0, 0);
    a.lr = lr;
    return launch_bpr_apply(h, a, d_users, d_pos, d_neg, n, (cudaStream_t)stream);
}

int bfl_sgd_sample_device(bfl_sgd_t* h, int64_t row_begin, int64_t row_end, int32_t* d_users, int32_t* d_pos,
                          int32_t* d_neg, void* stream) {
    if (!h || !h->factors_ready) BFL_FAIL(BFL_ERR_STATE, "factors not bound");
    if (!h->d_indptr || !h->d_keys) BFL_FAIL(BFL_ERR_STATE, "no device CSR bound");
    if (row_begin < 0 || row_end > h->csr_rows || row_end <= row_begin) BFL_FAIL(BFL_ERR_ARG, "bad row range");
    int64_t ends[2] = {0, 0};
    if (row_begin > 0)
        BFL_CUDA(cudaMemcpyAsync(&ends[0], h->d_indptr + row_begin - 1, sizeof(int64_t), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    BFL_CUDA(cudaMemcpyAsync(&ends[1], h->d_indptr + row_end - 1, sizeof(int64_t), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    BFL_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    SgdArgs a;
    fill_args(h, a, h->d_keys, 0, row_begin, row_end, ends[0], ends[1]);
    const int64_t cnt = (ends[1] - ends[0]) * h->num_neg;
    if (cnt <= 0) return BFL_OK;
    const int grid = (int)std::min<int64_t>((cnt + 255) / 256, (int64_t)h->num_sms * 32);
    bpr_sample_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(a, d_users, d_pos, d_neg);
    BFL_LAUNCHED();
    return BFL_OK;
}

float* bfl_sgd_grad_device(bfl_sgd_t* h, int which) {
    if (!h) return nullptr;
    return which == 0 ? h->gP.p : (which == 1 ? h->gQ.p : h->gQb.p);
}

int32_t* bfl_sgd_count_device(bfl_sgd_t* h, int which) {
    if (!h) return nullptr;
    return which == 0 ? h->cP.p : h->cQ.p;
}

int bfl_sgd_set_trace_device(bfl_sgd_t* h, int32_t* d_trials, int32_t* d_negs) {
    if (!h) BFL_FAIL(BFL_ERR_ARG, "null handle");
    h->trace_trials = d_trials;
    h->trace_negs = d_negs;
    return BFL_OK;
}

int bfl_sgd_epoch(bfl_sgd_t* h) { return h ? h->epoch : -1; }
double bfl_sgd_current_lr(bfl_sgd_t* h) { return h ? h->cur_lr : 0.0; }

int bfl_sgd_read_stats(bfl_sgd_t* h, double* loss_sum, int64_t* num_updates) {
    if (!h || !h->opt_set) BFL_FAIL(BFL_ERR_STATE, "init() first");
    double l = 0.0;
    unsigned long long u = 0;
    BFL_CUDA(cudaDeviceSynchronize());
    BFL_CUDA(cudaMemcpy(&l, h->d_stat.p, sizeof(double), cudaMemcpyDeviceToHost));
    BFL_CUDA(cudaMemcpy(&u, h->d_upd.p, sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    if (loss_sum) *loss_sum = l;
    if (num_updates) *num_updates = (int64_t)u;
    return BFL_OK;
}

}  // extern "C"
