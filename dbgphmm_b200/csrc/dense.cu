// dense.cu — dense DP rows over all N node states.
//
// One CTA computes one DP row of one chunk (DENSE_CORE consecutive relabelled nodes) for one read:
// the previous row of the chunk and of its 6-hop upstream halo is staged in shared memory, Match/Ins are
// computed for every local node within 5 hops, and the bounded Del chain (fd0 + 4 x fdt, forward.rs:423-466;
// bd0 + 4 x bdt, backward.rs:299-343) is propagated through shared memory with the valid region shrinking by
// one hop per round, so a row costs one read and one write of the 28 B cell per node.
// f_step: forward.rs:276-306 (fm :337, fi :378, fd0 :480, fdt :510, fib :541, fe :554)
// b_step: backward.rs:216-261 (bd0 :354, bdt :387, bm :423, bi :462, bmb :499, bib :535)
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include "dense.h"
#include "engine.h"

struct PlanView {
    const uint32_t *chunk_start, *loc_base, *loc_node, *nle, *le_off, *le_eid;
    const uint16_t* le_idx;
    const uint16_t* fp_idx; const uint32_t* fp_eid; const uint32_t* fx_off; const uint16_t* fx_idx; const uint32_t* fx_eid;
    const uint32_t* rl_node; const uint16_t* rl_par; const uint32_t* rl_eid; const uint8_t* rl_flag; const uint16_t* rl_core;
    const uint32_t* rx_off; const uint16_t* rx_idx; const uint32_t* rx_eid;
};
struct GraphView {
    uint32_t N, E;
    const uint8_t* emission;
    const double *init, *trans;
    const uint32_t* orig_of;
};
static PlanView plan_view(const DevPlan& p) { return PlanView{p.chunk_start, p.loc_base, p.loc_node, p.nle, p.le_off, p.le_eid, p.le_idx, p.fp_idx, p.fp_eid, p.fx_off, p.fx_idx, p.fx_eid,
                                                                   p.rl_node, p.rl_par, p.rl_eid, p.rl_flag, p.rl_core, p.rx_off, p.rx_idx, p.rx_eid}; }
static GraphView graph_view(const dbgphmm_model* m) { return GraphView{m->N, m->E, m->d_emission, m->d_init, m->d_trans, m->d_orig_of}; }

__device__ __forceinline__ uint64_t slab_of(const DJob& jb, uint32_t s) { return jb.slab0 + (jb.slab_mod ? (s % jb.slab_mod) : s); }

__device__ __forceinline__ XF warp_xsum(XF a) {
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        XF b;
        b.v = __shfl_down_sync(0xffffffffu, a.v, o);
        b.e = __shfl_down_sync(0xffffffffu, a.e, o);
        a = xadd(a, b);
    }
    return a;
}
__device__ __forceinline__ XF warp_xmax(XF a) {
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        XF b;
        b.v = __shfl_down_sync(0xffffffffu, a.v, o);
        b.e = __shfl_down_sync(0xffffffffu, a.e, o);
        if (xgt(b, a)) a = b;
    }
    return a;
}

#define SMEM_CARVE()                                                   \
    extern __shared__ __align__(16) unsigned char smem_raw[];          \
    double* pm = (double*)smem_raw;                                    \
    double* pi = pm + DENSE_LMAX;                                      \
    double* pd = pi + DENSE_LMAX;                                      \
    double* cm = pd + DENSE_LMAX;                                      \
    double* ci = cm + DENSE_LMAX;                                      \
    double* dv0 = ci + DENSE_LMAX;                                     \
    double* dv1 = dv0 + DENSE_LMAX;                                    \
    int* pex = (int*)(dv1 + DENSE_LMAX);                               \
    int* cex = pex + DENSE_LMAX;                                       \
    int* de0 = cex + DENSE_LMAX;                                       \
    int* de1 = de0 + DENSE_LMAX;
#define DENSE_SMEM_BYTES (DENSE_LMAX * (7 * 8 + 4 * 4))

// ------------------------------------------------------------------------------------------------ forward
__global__ void __launch_bounds__(DENSE_THREADS)
k_dense_fwd(PlanView P, GraphView G, LinParams lp, const DJob* __restrict__ jobs, uint32_t s, const uint8_t* __restrict__ bases,
            const RowDesc* __restrict__ desc, const int* __restrict__ active, char* __restrict__ pool, uint64_t slab_bytes,
            uint32_t Np, XF* __restrict__ partials, uint32_t n_chunks, const unsigned long long* __restrict__ worklist) {
  // exact path: runs over the (job, chunk) tiles the common-frame kernel could not take (worklist[0] = count)
  const unsigned long long n_work = worklist[0];
  for (unsigned long long wi = blockIdx.x; wi < n_work; wi += gridDim.x) {
    const unsigned long long wl = worklist[1 + wi];
    const uint32_t job_idx = (uint32_t)(wl >> 32);
    const DJob jb = jobs[job_idx];
    SMEM_CARVE();
    const int tid = threadIdx.x;
    const uint32_t c = (uint32_t)wl;
    __syncthreads();
    const int row = jb.first_row + (int)s;
    const uint8_t x = bases[jb.base_off + row];
    const uint32_t lb = P.loc_base[c], lo = lb + c;
    const uint32_t* nle = P.nle + (size_t)c * 8;
    const int n0 = nle[0], n4 = nle[4], n5 = nle[5], nL = nle[6];
    const double* init = G.init + (size_t)jb.x * G.N;
    const double* trans = G.trans + (size_t)jb.x * G.E;
    // begin-state scalars of the previous row (f_init: mb = 1, forward.rs:255-266)
    XF mbp, ibp;
    if (row == 0) { mbp = xf(1.0, 0); ibp = xf_zero(); }
    else { mbp = desc[jb.desc0 + row - 1].mb; ibp = desc[jb.desc0 + row - 1].ib; }
    const XF ib_cur = xmul(xadd(xmul(mbp, lp.p_MI), xmul(ibp, lp.p_II)), lp.p_random);  // fib, forward.rs:541-545
    const XF fb0 = xadd(xmul(mbp, lp.p_MM), xmul(ibp, lp.p_IM));                        // begin part of fm
    // ---- stage the previous row of the local set
    const int pk = (s == 0) ? jb.prev0_kind : PREV_SLAB;
    if (pk == PREV_SLAB) {
        const char* sl = pool + (s == 0 ? jb.prev0_slab : slab_of(jb, s - 1)) * slab_bytes;
        const double* gm = (const double*)sl; const double* gi = gm + Np; const double* gd = gi + Np; const int* ge = (const int*)(gd + Np);
        for (int j = tid; j < nL; j += DENSE_THREADS) {
            uint32_t g = P.loc_node[lb + j];
            pm[j] = gm[g]; pi[j] = gi[g]; pd[j] = gd[g]; pex[j] = ge[g];
        }
    } else {
        for (int j = tid; j < nL; j += DENSE_THREADS) { pm[j] = 0.0; pi[j] = 0.0; pd[j] = 0.0; pex[j] = 0; }
    }
    __syncthreads();
    // ---- round A: Match / Ins of the current row (fm, fi)
    for (int j = tid; j < n5; j += DENSE_THREADS) {
        uint32_t g = P.loc_node[lb + j];
        XF acc = xf_zero();
        for (uint32_t a = P.le_off[lo + j], ae = P.le_off[lo + j + 1]; a < ae; a++) {
            int l = P.le_idx[a];
            double t = trans[P.le_eid[a]];
            acc = xadd(acc, xf(t * (lp.p_MM * pm[l] + lp.p_IM * pi[l] + lp.p_DM * pd[l]), pex[l]));
        }
        acc = xadd(acc, xmul(fb0, init[g]));
        XF m = xmul(acc, G.emission[g] == x ? lp.p_match : lp.p_mismatch);
        XF i = xf(lp.p_random * (lp.p_MI * pm[j] + lp.p_II * pi[j] + lp.p_DI * pd[j]), pex[j]);
        int Em = xexp(m), Ei = xexp(i), Ec = Em > Ei ? Em : Ei;
        if (Ec == XF_ZERO_E) { cm[j] = 0.0; ci[j] = 0.0; cex[j] = 0; }
        else {
            cm[j] = m.v == 0.0 ? 0.0 : m.v * pow2i(m.e - Ec);
            ci[j] = i.v == 0.0 ? 0.0 : i.v * pow2i(i.e - Ec);
            cex[j] = Ec;
        }
    }
    __syncthreads();
    // ---- round B: fd0 ; rounds C1..C4: fdt.  d accumulates in registers (slot q of this thread).
    XF dacc[DENSE_SLOTS];
#pragma unroll
    for (int q = 0; q < DENSE_SLOTS; q++) {
        dacc[q] = xf_zero();
        int j = tid + q * DENSE_THREADS;
        if (j < n4) {
            uint32_t g = P.loc_node[lb + j];
            XF acc = xf_zero();
            for (uint32_t a = P.le_off[lo + j], ae = P.le_off[lo + j + 1]; a < ae; a++) {
                int l = P.le_idx[a];
                double t = trans[P.le_eid[a]];
                acc = xadd(acc, xf(t * (lp.p_MD * cm[l] + lp.p_ID * ci[l]), cex[l]));
            }
            acc = xadd(acc, xmul(ib_cur, lp.p_ID * init[g]));  // mb of the current row is 0 (fmb, forward.rs:531-533)
            dv0[j] = acc.v; de0[j] = acc.e;
            dacc[q] = acc;
        }
    }
    __syncthreads();
    double* dprev = dv0; int* eprev = de0; double* dcur = dv1; int* ecur = de1;
#pragma unroll 1
    for (int t = 1; t < N_DEL_ROUNDS; t++) {
        const int nb = nle[4 - t];
#pragma unroll
        for (int q = 0; q < DENSE_SLOTS; q++) {
            int j = tid + q * DENSE_THREADS;
            if (j < nb) {
                XF acc = xf_zero();
                for (uint32_t a = P.le_off[lo + j], ae = P.le_off[lo + j + 1]; a < ae; a++) {
                    int l = P.le_idx[a];
                    double tr = trans[P.le_eid[a]];
                    acc = xadd(acc, xf(tr * lp.p_DD * dprev[l], eprev[l]));
                }
                dcur[j] = acc.v; ecur[j] = acc.e;
                dacc[q] = xadd(dacc[q], acc);
            }
        }
        __syncthreads();
        double* tv = dprev; dprev = dcur; dcur = tv;
        int* te = eprev; eprev = ecur; ecur = te;
    }
    // ---- store the chunk's cells, reduce sum(m+i+d) for fe (forward.rs:554-558)
    char* so = pool + slab_of(jb, s) * slab_bytes;
    double* om = (double*)so; double* oi = om + Np; double* od = oi + Np; int* oe = (int*)(od + Np);
    const uint32_t g0 = P.chunk_start[c];
    XF part = xf_zero();
#pragma unroll
    for (int q = 0; q < DENSE_SLOTS; q++) {
        int j = tid + q * DENSE_THREADS;
        if (j < n0) {
            Cell cl = cell_pack(xf(cm[j], cex[j]), xf(ci[j], cex[j]), dacc[q]);
            om[g0 + j] = cl.m; oi[g0 + j] = cl.i; od[g0 + j] = cl.d; oe[g0 + j] = cl.e;
            part = xadd(part, xf(cl.m + cl.i + cl.d, cl.e));
        }
    }
    part = warp_xsum(part);
    __shared__ XF wsum[DENSE_THREADS / 32];
    if ((tid & 31) == 0) wsum[tid >> 5] = part;
    __syncthreads();
    if (tid == 0) {
        XF tot = xf_zero();
        for (int w = 0; w < DENSE_THREADS / 32; w++) tot = xadd(tot, wsum[w]);
        partials[(size_t)job_idx * n_chunks + c] = tot;
    }
  }
}

// row reduction of the forward step: e = p_end * sum ; ib ; mb = 0
__global__ void k_dense_fwd_finish(LinParams lp, const DJob* __restrict__ jobs, uint32_t s, RowDesc* __restrict__ desc,
                                   const int* __restrict__ active, const XF* __restrict__ partials, uint32_t n_chunks, int pair_slabs = 0) {
    const DJob jb = jobs[blockIdx.x];
    if (s >= jb.n_steps) return;
    if (jb.active_idx >= 0 && !active[jb.active_idx]) return;
    __shared__ XF sh[256];
    XF acc = xf_zero();
    for (uint32_t c = threadIdx.x; c < n_chunks; c += blockDim.x) acc = xadd(acc, partials[(size_t)blockIdx.x * n_chunks + c]);
    sh[threadIdx.x] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        XF tot = xf_zero();
        for (int t = 0; t < (int)blockDim.x; t++) tot = xadd(tot, sh[t]);
        int row = jb.first_row + (int)s;
        XF mbp, ibp;
        if (row == 0) { mbp = xf(1.0, 0); ibp = xf_zero(); }
        else { mbp = desc[jb.desc0 + row - 1].mb; ibp = desc[jb.desc0 + row - 1].ib; }
        RowDesc r;
        r.kind = ROW_DENSE; r.n_ent = 0; r.n_mi = 0; r.n_d = 0;
        r.off = pair_slabs ? jb.slab0 + ((s >> 1) & 1) : jb.slab0 + (jb.slab_mod ? (s % jb.slab_mod) : s);
        r.mb = xf_zero();
        r.ib = xnorm(xmul(xadd(xmul(mbp, lp.p_MI), xmul(ibp, lp.p_II)), lp.p_random));
        r.e = xnorm(xmul(tot, lp.p_end));
        desc[jb.desc0 + row] = r;
    }
}

// ------------------------------------------------------------------------------------------------ backward
__global__ void __launch_bounds__(DENSE_THREADS)
k_dense_bwd(PlanView P, GraphView G, LinParams lp, const DJob* __restrict__ jobs, uint32_t s, const uint8_t* __restrict__ bases,
            const int* __restrict__ active, char* __restrict__ pool, uint64_t slab_bytes, uint32_t Np, XF* __restrict__ partials,
            uint32_t n_chunks, const unsigned long long* __restrict__ worklist) {
  const unsigned long long n_work = worklist[0];
  for (unsigned long long wi = blockIdx.x; wi < n_work; wi += gridDim.x) {
    const unsigned long long wl = worklist[1 + wi];
    const uint32_t job_idx = (uint32_t)(wl >> 32);
    const DJob jb = jobs[job_idx];
    SMEM_CARVE();
    (void)pd; (void)ci;
    const int tid = threadIdx.x;
    const uint32_t c = (uint32_t)wl;
    __syncthreads();
    const int row = jb.first_row - (int)s;
    const uint8_t x = bases[jb.base_off + row];
    const uint32_t lb = P.loc_base[c], lo = lb + c;
    const uint32_t* nle = P.nle + (size_t)c * 8;
    const int n0 = nle[0], n1 = nle[1], n5 = nle[5], nL = nle[6];
    const double* init = G.init + (size_t)jb.x * G.N;
    const double* trans = G.trans + (size_t)jb.x * G.E;
    // ---- stage the next row (i+1): m'', i'' pre-multiplied views are formed on the fly
    const int pk = (s == 0) ? jb.prev0_kind : PREV_SLAB;
    if (pk == PREV_SLAB) {
        const char* sl = pool + (s == 0 ? jb.prev0_slab : slab_of(jb, s - 1)) * slab_bytes;
        const double* gm = (const double*)sl; const double* gi = gm + Np; const int* ge = (const int*)(gi + 2 * (size_t)Np);
        for (int j = tid; j < nL; j += DENSE_THREADS) {
            uint32_t g = P.loc_node[lb + j];
            // cm holds e_l(x) * m''[l] (the emission of the child is what every use multiplies by)
            double em = G.emission[g] == x ? lp.p_match : lp.p_mismatch;
            pm[j] = gm[g] * em; pi[j] = gi[g]; pex[j] = ge[g];
        }
    } else {  // b_init: m = i = d = p_end (backward.rs:197-211)
        for (int j = tid; j < nL; j += DENSE_THREADS) {
            uint32_t g = P.loc_node[lb + j];
            double em = G.emission[g] == x ? lp.p_match : lp.p_mismatch;
            pm[j] = lp.p_end * em; pi[j] = lp.p_end; pex[j] = 0;
        }
    }
    __syncthreads();
    // ---- bd0 for depth <= 5, then bdt with the valid region shrinking; d of depth <= 1 ends in (cm, cex)
    XF dacc[DENSE_SLOTS];
#pragma unroll
    for (int q = 0; q < DENSE_SLOTS; q++) {
        dacc[q] = xf_zero();
        int j = tid + q * DENSE_THREADS;
        if (j < n5) {
            XF acc = xf_zero();
            for (uint32_t a = P.le_off[lo + j], ae = P.le_off[lo + j + 1]; a < ae; a++) {
                int l = P.le_idx[a];
                double t = trans[P.le_eid[a]];
                acc = xadd(acc, xf(t * lp.p_DM * pm[l], pex[l]));
            }
            acc = xadd(acc, xf(lp.p_DI * lp.p_random * pi[j], pex[j]));
            dv0[j] = acc.v; de0[j] = acc.e;
            dacc[q] = acc;
        }
    }
    __syncthreads();
    double* dprev = dv0; int* eprev = de0; double* dcur = dv1; int* ecur = de1;
#pragma unroll 1
    for (int t = 1; t < N_DEL_ROUNDS; t++) {
        const int nb = nle[5 - t];
#pragma unroll
        for (int q = 0; q < DENSE_SLOTS; q++) {
            int j = tid + q * DENSE_THREADS;
            if (j < nb) {
                XF acc = xf_zero();
                for (uint32_t a = P.le_off[lo + j], ae = P.le_off[lo + j + 1]; a < ae; a++) {
                    int l = P.le_idx[a];
                    double tr = trans[P.le_eid[a]];
                    acc = xadd(acc, xf(tr * lp.p_DD * dprev[l], eprev[l]));
                }
                dcur[j] = acc.v; ecur[j] = acc.e;
                dacc[q] = xadd(dacc[q], acc);
            }
        }
        __syncthreads();
        double* tv = dprev; dprev = dcur; dcur = tv;
        int* te = eprev; eprev = ecur; ecur = te;
    }
#pragma unroll
    for (int q = 0; q < DENSE_SLOTS; q++) {
        int j = tid + q * DENSE_THREADS;
        if (j < n1) { cm[j] = dacc[q].v; cex[j] = dacc[q].e; }
    }
    __syncthreads();
    // ---- bm, bi of the chunk; partial sums of bmb / bib over the chunk's nodes
    char* so = pool + slab_of(jb, s) * slab_bytes;
    double* om = (double*)so; double* oi = om + Np; double* od = oi + Np; int* oe = (int*)(od + Np);
    const uint32_t g0 = P.chunk_start[c];
    XF pmb = xf_zero(), pib = xf_zero();
#pragma unroll
    for (int q = 0; q < DENSE_SLOTS; q++) {
        int j = tid + q * DENSE_THREADS;
        if (j < n0) {
            XF am = xf_zero(), ai = xf_zero();
            for (uint32_t a = P.le_off[lo + j], ae = P.le_off[lo + j + 1]; a < ae; a++) {
                int l = P.le_idx[a];
                double t = trans[P.le_eid[a]];
                XF tm = xf(t * pm[l], pex[l]);        // t * e_l(x) * m''[l]
                XF td = xf(t * cm[l], cex[l]);        // t * d[l]
                am = xadd(am, xadd(xmul(tm, lp.p_MM), xmul(td, lp.p_MD)));
                ai = xadd(ai, xadd(xmul(tm, lp.p_IM), xmul(td, lp.p_ID)));
            }
            am = xadd(am, xf(lp.p_MI * lp.p_random * pi[j], pex[j]));
            ai = xadd(ai, xf(lp.p_II * lp.p_random * pi[j], pex[j]));
            Cell cl = cell_pack(am, ai, dacc[q]);
            uint32_t g = g0 + j;
            om[g] = cl.m; oi[g] = cl.i; od[g] = cl.d; oe[g] = cl.e;
            // bmb / bib terms of this node (backward.rs:499-555)
            double in = init[g];
            XF um = xf(pm[j], pex[j]);
            pmb = xadd(pmb, xmul(xadd(xmul(um, lp.p_MM), xmul(dacc[q], lp.p_MD)), in));
            pib = xadd(pib, xmul(xadd(xmul(um, lp.p_IM), xmul(dacc[q], lp.p_ID)), in));
        }
    }
    pmb = warp_xsum(pmb); pib = warp_xsum(pib);
    __shared__ XF wsum[2][DENSE_THREADS / 32];
    if ((tid & 31) == 0) { wsum[0][tid >> 5] = pmb; wsum[1][tid >> 5] = pib; }
    __syncthreads();
    if (tid == 0) {
        XF a = xf_zero(), b = xf_zero();
        for (int w = 0; w < DENSE_THREADS / 32; w++) { a = xadd(a, wsum[0][w]); b = xadd(b, wsum[1][w]); }
        partials[((size_t)job_idx * n_chunks + c) * 2] = a;
        partials[((size_t)job_idx * n_chunks + c) * 2 + 1] = b;
    }
  }
}

__global__ void k_dense_bwd_finish(LinParams lp, const DJob* __restrict__ jobs, uint32_t s, RowDesc* __restrict__ desc,
                                   const int* __restrict__ active, const XF* __restrict__ partials, uint32_t n_chunks, int pair_slabs = 0) {
    const DJob jb = jobs[blockIdx.x];
    if (s >= jb.n_steps) return;
    if (jb.active_idx >= 0 && !active[jb.active_idx]) return;
    __shared__ XF sh[2][256];
    XF a = xf_zero(), b = xf_zero();
    for (uint32_t c = threadIdx.x; c < n_chunks; c += blockDim.x) {
        a = xadd(a, partials[((size_t)blockIdx.x * n_chunks + c) * 2]);
        b = xadd(b, partials[((size_t)blockIdx.x * n_chunks + c) * 2 + 1]);
    }
    sh[0][threadIdx.x] = a; sh[1][threadIdx.x] = b;
    __syncthreads();
    if (threadIdx.x == 0) {
        XF ta = xf_zero(), tb = xf_zero();
        for (int t = 0; t < (int)blockDim.x; t++) { ta = xadd(ta, sh[0][t]); tb = xadd(tb, sh[1][t]); }
        int row = jb.first_row - (int)s;
        XF ibn = (row == (int)jb.len - 1) ? xf_zero() : desc[jb.desc0 + row + 1].ib;  // b_init: ib = 0
        RowDesc r;
        r.kind = ROW_DENSE; r.n_ent = 0; r.n_mi = 0; r.n_d = 0;
        r.off = pair_slabs ? jb.slab0 + ((s >> 1) & 1) : jb.slab0 + (jb.slab_mod ? (s % jb.slab_mod) : s);
        r.mb = xnorm(xadd(ta, xmul(ibn, lp.p_MI * lp.p_random)));
        r.ib = xnorm(xadd(tb, xmul(ibn, lp.p_II * lp.p_random)));
        r.e = xf_zero();
        desc[jb.desc0 + row] = r;
    }
}

#define EXP_NONE_LO_ 0x3fffffff
#define EXP_NONE_HI_ (-0x3fffffff)
// ------------------------------------------------------------------------------------------------ common-frame kernels
// Fast path of the dense row step, one WARP per tile (DENSE_CORE consecutive nodes + halo), no block barriers.  When every non-zero
// input of a tile (previous row of the tile and its halo, begin-state terms) lies within `span` binades of the largest one, all of
// them are brought to ONE exponent E_ref with exact power-of-two scalings and the step runs in plain f64 (no operand can underflow
// inside the step: `span` leaves room for every factor a row applies).  Tiles that do not qualify go to the exact kernel above.
// Register-stencil form of that step (one warp per tile, one exponent frame per tile, exact scalings): every lane OWNS
// RS_PER_LANE consecutive positions of the tile's register layout (model.cu: [chain above the tile head | core | rest of
// the halo in chains]).  For almost every node the first upstream neighbour is the previous position, i.e. the lane's own
// previous register or one shuffle from the lane below, so Match/Ins/Del need no shared-memory gathers; nodes whose
// neighbour is elsewhere (bit 0 of rl_flag) or that have further upstream edges (bit 1) take an out-of-line path through
// the position-indexed copies kept in shared memory.  All positions are computed in every round: cells whose inputs lie
// outside the tile become garbage-but-finite and can reach the core only after more than 6 hops, i.e. never.
// Per-(job, step) scalars of the fast kernel, written by k_dense_prep before every step so that the tile warps get them
// (and the address of the previous row, which the row prefetch needs one job ahead) with one 64-byte asynchronous copy.
struct __align__(16) JStep {
    unsigned long long prev_ptr, out_ptr;   // slab of the previous row / of this row
    XF fb0, ib_cur;                         // forward: begin part of fm, fib (forward.rs:541-545)
    int valid, x, pk; unsigned int base;    // 0 = skip ; parameter set ; PREV_* ; read base
};
static_assert(sizeof(JStep) == 64, "JStep is copied as four 16-byte pieces");

template <bool FWD>
__global__ void k_dense_prep(LinParams lp, const DJob* __restrict__ jobs, uint32_t n_jobs, uint32_t s, const uint8_t* __restrict__ bases,
                             const RowDesc* __restrict__ desc, const int* __restrict__ active, char* __restrict__ pool, uint64_t slab_bytes,
                             JStep* __restrict__ out) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_jobs) return;
    const DJob jb = jobs[j];
    JStep js;
    js.prev_ptr = 0; js.out_ptr = 0; js.fb0 = xf_zero(); js.ib_cur = xf_zero(); js.valid = 0; js.x = 0; js.pk = 0; js.base = 0;
    if (s < jb.n_steps && !(jb.active_idx >= 0 && !active[jb.active_idx])) {
        const int row = FWD ? jb.first_row + (int)s : jb.first_row - (int)s;
        js.valid = 1; js.x = (int)jb.x; js.base = bases[jb.base_off + row];
        js.pk = (s == 0) ? jb.prev0_kind : PREV_SLAB;
        js.prev_ptr = (unsigned long long)(pool + (s == 0 ? jb.prev0_slab : slab_of(jb, s - 1)) * slab_bytes);
        js.out_ptr = (unsigned long long)(pool + slab_of(jb, s) * slab_bytes);
        if (FWD) {
            XF mbp, ibp;
            if (row == 0) { mbp = xf(1.0, 0); ibp = xf_zero(); }
            else { mbp = desc[jb.desc0 + row - 1].mb; ibp = desc[jb.desc0 + row - 1].ib; }
            js.ib_cur = xmul(xadd(xmul(mbp, lp.p_MI), xmul(ibp, lp.p_II)), lp.p_random);
            js.fb0 = xadd(xmul(mbp, lp.p_MM), xmul(ibp, lp.p_IM));
        }
    }
    out[j] = js;
}

__device__ __forceinline__ void cp_async4(void* smem, const void* gmem) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async8(void* smem, const void* gmem) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

#define RS_PER_LANE (DENSE_LMAX / 32)
// per warp: frame-scaled copies sa, sb, sc + se ; raw prefetched previous row rm, ri, rd + re ; two JStep slots ;
// init of the owned positions and node ids of the IO positions (lane-private, kept out of the register file)
#define RS_SMEM_PER_WARP (2 * DENSE_LMAX * (3 * 8 + 4) + 2 * 64 + DENSE_LMAX * (8 + 4))
#define RS_SMEM_BYTES (WT_WARPS * RS_SMEM_PER_WARP)

__device__ __forceinline__ double rs_extras(const uint32_t* __restrict__ rx_off, const uint16_t* __restrict__ rx_idx, const uint32_t* __restrict__ rx_eid,
                                         const double* __restrict__ trans, uint32_t xidx, const double* a, const double* b, const double* c,
                                         double ca, double cb, double cc) {
    double acc = 0.0;
    for (uint32_t e = rx_off[xidx], ee = rx_off[xidx + 1]; e < ee; e++) {
        const int l = rx_idx[e];
        acc += trans[rx_eid[e]] * (ca * a[l] + cb * b[l] + cc * c[l]);
    }
    return acc;
}
// slot 0 of a lane with several extra upstream edges: out-of-line loop (rare)
#define RS_EXTRAS0(A, B, C, ca, cb, cc) rs_extras(P.rx_off, P.rx_idx, P.rx_eid, trans, (size_t)c * (DENSE_LMAX + 1) + RS_PER_LANE * lane, (A), (B), (C), (ca), (cb), (cc))
// first upstream neighbour of slot k: slot 0 reads the staged copy at pp0, the others the previous register
#define RS_UP(arr, buf, k) ((k) == 0 ? (buf)[pp0] : arr[(k) > 0 ? (k) - 1 : 0])
// publish the owned values other lanes read (first-neighbour / extra sources): normally just the last slot
#define RS_STAGE(buf, arr) do {                                                                                   \
        if (tile_plain) (buf)[RS_PER_LANE * lane + RS_PER_LANE - 1] = arr[RS_PER_LANE - 1];                      \
        else { _Pragma("unroll") for (int k_ = 0; k_ < RS_PER_LANE; k_++) if ((srcm >> k_) & 1) (buf)[RS_PER_LANE * lane + k_] = arr[k_]; } \
    } while (0)

template <bool FWD>
__global__ void __launch_bounds__(WT_WARPS * 32, WT_MIN_CTAS)
k_dense_reg(PlanView P, GraphView G, LinParams lp, const JStep* __restrict__ jstep, uint32_t n_jobs, uint32_t Np,
            XF* __restrict__ partials, uint32_t n_tiles, int span, unsigned long long* __restrict__ worklist, uint32_t jpc) {
    extern __shared__ __align__(16) unsigned char rs_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t c = blockIdx.x * WT_WARPS + warp;
    if (c >= n_tiles) return;   // no block-wide barrier anywhere in this kernel
    double* sa = (double*)(rs_raw + (size_t)warp * RS_SMEM_PER_WARP);
    double* sb = sa + DENSE_LMAX; double* sc = sb + DENSE_LMAX; int* se = (int*)(sc + DENSE_LMAX);
    double* rm = (double*)(se + DENSE_LMAX); double* ri = rm + DENSE_LMAX; double* rd = ri + DENSE_LMAX; int* re = (int*)(rd + DENSE_LMAX);
    JStep* sj = (JStep*)(re + DENSE_LMAX);
    double* sinit = (double*)(sj + 2);               // init of the owned positions, [slot][lane] (lane-private, conflict-free)
    uint32_t* sion = (uint32_t*)(sinit + DENSE_LMAX); // node of IO position 32 q + lane, [q][lane] (lane-private)
    const uint32_t tbase = c * DENSE_LMAX;
    const uint32_t g0 = P.chunk_start[c], ncore = P.chunk_start[c + 1] - g0;
    // ---- tile structure: slot-major node ids for global IO, owned positions for the recurrences
    unsigned int emc = 0;     // 3-bit base codes of the owned positions
    unsigned int cmask = 0;   // bit k: owned position k is a core node ; bit 8 + q: IO position 32 q + lane is a core node
    unsigned int srcm = 0;    // bit k: owned position k is read by another position's slot 0 or as an extra source
#pragma unroll
    for (int q = 0; q < RS_PER_LANE; q++) {
        const uint32_t nd = P.rl_node[tbase + 32 * q + lane];
        sion[32 * q + lane] = nd;
        cmask |= (nd - g0 < ncore ? 1u : 0u) << (8 + q);
        if (nd != 0xffffffffu) cmask |= 1u << (16 + q);   // bit 16 + q: IO position holds a node
    }
#pragma unroll
    for (int k = 0; k < RS_PER_LANE; k++) {
        const uint32_t nd = P.rl_node[tbase + RS_PER_LANE * lane + k];
        unsigned int code = 4;
        if (nd != 0xffffffffu) { unsigned char ch = G.emission[nd]; code = ch == 'n' ? 4 : ((ch >> 1) & 3); }
        emc |= code << (3 * k);
        cmask |= (nd - g0 < ncore ? 1u : 0u) << k;
        srcm |= (unsigned int)((P.rl_flag[tbase + RS_PER_LANE * lane + k] >> 2) & 1) << k;
    }
    const bool tile_plain = !__any_sync(0xffffffffu, (srcm & ((1u << (RS_PER_LANE - 1)) - 1)) != 0);
    // Slot 0 of the lane: position of its first upstream neighbour (the layout puts every node whose neighbour is not the
    // previous position, or that has further upstream edges, on a slot 0).  ONE extra edge is kept in registers
    // (source position xp, edge id xe -> transition xt); more take the out-of-line loop.
    const int pp0 = P.rl_par[tbase + RS_PER_LANE * lane];
    int xp = pp0; uint32_t xe = 0xffffffffu; double xt = 0.0; bool ext0 = false;   // (lanes without an extra read a published, finite value times 0)
    if (P.rl_flag[tbase + RS_PER_LANE * lane] & 2) {
        const size_t xi = (size_t)c * (DENSE_LMAX + 1) + RS_PER_LANE * lane;
        const uint32_t a0 = P.rx_off[xi], a1 = P.rx_off[xi + 1];
        if (a1 - a0 == 1) { xp = P.rx_idx[a0]; xe = P.rx_eid[a0]; } else ext0 = true;
    }
    const bool tile_has_x = __any_sync(0xffffffffu, xe != 0xffffffffu);
    const bool tile_has_xx = __any_sync(0xffffffffu, ext0);
    double tr[RS_PER_LANE];
#pragma unroll
    for (int k = 0; k < RS_PER_LANE; k++) tr[k] = 0.0;
    int staged_x = -1;
    const double* trans = G.trans;
    const uint32_t job0 = blockIdx.y * jpc;
    const uint32_t n_here = job0 < n_jobs ? min(jpc, n_jobs - job0) : 0u;
    // Software pipeline over the jobs of this warp: the scalars of job j + 2 and the previous-row tile of job j + 1 are
    // in flight (cp.async into shared memory, no registers held) while job j is computed.
    auto fetch_js = [&](uint32_t jj, int slot) {
        if (lane < 4) cp_async16((char*)&sj[slot] + 16 * lane, (const char*)&jstep[job0 + jj] + 16 * lane);
    };
    auto fetch_rows = [&](const JStep& js) {
        if (!js.valid || js.pk != PREV_SLAB) return;
        const double* gm = (const double*)js.prev_ptr; const double* gi = gm + Np; const double* gd = gi + Np; const int* ge = (const int*)(gd + Np);
#pragma unroll
        for (int q = 0; q < RS_PER_LANE; q++) {
            if ((cmask >> (16 + q)) & 1) {
                const int pos = 32 * q + lane; const uint32_t g = sion[pos];
                cp_async8(&rm[pos], &gm[g]); cp_async8(&ri[pos], &gi[g]); if (FWD) cp_async8(&rd[pos], &gd[g]); cp_async4(&re[pos], &ge[g]);
            }
        }
    };
    if (n_here == 0) return;
#pragma unroll
    for (int q = 0; q < RS_PER_LANE; q++) { const int pos = 32 * q + lane; rm[pos] = 0.0; ri[pos] = 0.0; rd[pos] = 0.0; re[pos] = 0; }   // pads stay zero
    fetch_js(0, 0);
    if (n_here > 1) fetch_js(1, 1);
    cp_async_wait_all();
    __syncwarp();
    fetch_rows(sj[0]);
    for (uint32_t jj = 0; jj < n_here; jj++) {
        const uint32_t job_idx = job0 + jj;
        const int slot = jj & 1;
        cp_async_wait_all();
        __syncwarp();
        // refill the raw row and this JStep slot for the jobs ahead ; call once every lane is done reading them
        auto prefetch_next = [&]() {
            __syncwarp();
            if (jj + 1 < n_here) fetch_rows(sj[slot ^ 1]);
            if (jj + 2 < n_here) fetch_js(jj + 2, slot);
        };
        // ---- per-job scalars and the prefetched previous row
        const int valid = sj[slot].valid, hx = sj[slot].x, pk = sj[slot].pk;
        const unsigned int xbase = sj[slot].base;
        const unsigned long long out_ptr = sj[slot].out_ptr;
        XF fb0 = xf_zero(), ib_cur = xf_zero();
        if (FWD) { fb0 = sj[slot].fb0; ib_cur = sj[slot].ib_cur; }
        if (!valid) { prefetch_next(); continue; }
        const unsigned int x = (xbase >> 1) & 3;   // base code of the read base
        if (hx != staged_x) {
            const double* init = G.init + (size_t)hx * G.N;
            trans = G.trans + (size_t)hx * G.E;
#pragma unroll
            for (int k = 0; k < RS_PER_LANE; k++) {
                const uint32_t nd = P.rl_node[tbase + RS_PER_LANE * lane + k], ed = P.rl_eid[tbase + RS_PER_LANE * lane + k];
                sinit[32 * k + lane] = nd == 0xffffffffu ? 0.0 : init[nd];
                tr[k] = ed == 0xffffffffu ? 0.0 : trans[ed];
            }
            xt = xe == 0xffffffffu ? 0.0 : trans[xe];
            staged_x = hx;
        }
        // ---- previous row: the prefetch buffer (filled slot-major by cp.async) is read position-major, so it doubles as
        // the transpose buffer; the first row of a read (constant previous row) is materialised in it
        if (pk != PREV_SLAB) {
            const double v0 = pk == PREV_B_INIT ? lp.p_end : 0.0;
#pragma unroll
            for (int q = 0; q < RS_PER_LANE; q++)
                if ((cmask >> (16 + q)) & 1) { const int pos = 32 * q + lane; rm[pos] = v0; ri[pos] = v0; rd[pos] = 0.0; re[pos] = 0; }
            __syncwarp();
        }
        double pm[RS_PER_LANE], pi[RS_PER_LANE], pd[RS_PER_LANE]; int pe[RS_PER_LANE];
        int elo = EXP_NONE_LO_, ehi = EXP_NONE_HI_;
#pragma unroll
        for (int k = 0; k < RS_PER_LANE; k++) {
            const int pos = RS_PER_LANE * lane + k;
            pm[k] = rm[pos]; pi[k] = ri[pos]; pd[k] = FWD ? rd[pos] : 0.0; pe[k] = re[pos];
            if (pm[k] + pi[k] + pd[k] != 0.0) { elo = pe[k] < elo ? pe[k] : elo; ehi = pe[k] > ehi ? pe[k] : ehi; }
        }
        if (FWD) {
            if (fb0.v != 0.0) { int e = xexp(fb0); elo = e < elo ? e : elo; ehi = e > ehi ? e : ehi; }
            if (ib_cur.v != 0.0) { int e = xexp(ib_cur); elo = e < elo ? e : elo; ehi = e > ehi ? e : ehi; }
        }
        elo = __reduce_min_sync(0xffffffffu, elo); ehi = __reduce_max_sync(0xffffffffu, ehi);
        double* om = (double*)out_ptr; double* oi = om + Np; double* od = oi + Np; int* oe = (int*)(od + Np);
        if (ehi == EXP_NONE_HI_) {  // nothing but zeros flows into this tile: the row is zero here
            prefetch_next();
            for (uint32_t j = lane; j < ncore; j += 32) { om[g0 + j] = 0.0; oi[g0 + j] = 0.0; od[g0 + j] = 0.0; oe[g0 + j] = 0; }
            if (lane == 0) {
                if (FWD) partials[(size_t)job_idx * n_tiles + c] = xf_zero();
                else { partials[((size_t)job_idx * n_tiles + c) * 2] = xf_zero(); partials[((size_t)job_idx * n_tiles + c) * 2 + 1] = xf_zero(); }
            }
            continue;
        }
        if (ehi - elo > span) {    // exponent range too wide for one frame: the exact kernel takes this tile
            prefetch_next();
            if (lane == 0) { unsigned long long w = atomicAdd(worklist, 1ull); worklist[1 + w] = ((unsigned long long)job_idx << 32) | c; }
            continue;
        }
        const int Eref = ehi;
        // ---- into the frame: the owned cells, the first upstream neighbour of slot 0 and the source of the register-held extra
        double u0m, u0i, u0d, x0m = 0.0, x0i = 0.0, x0d = 0.0;
        { const double s0 = pow2i(re[pp0] - Eref); u0m = rm[pp0] * s0; u0i = ri[pp0] * s0; u0d = FWD ? rd[pp0] * s0 : 0.0; }
        if (tile_has_x) { const double sx = pow2i(re[xp] - Eref); x0m = rm[xp] * sx; x0i = ri[xp] * sx; x0d = FWD ? rd[xp] * sx : 0.0; }
#pragma unroll
        for (int k = 0; k < RS_PER_LANE; k++) { const double sc_ = pow2i(pe[k] - Eref); pm[k] *= sc_; pi[k] *= sc_; pd[k] *= sc_; }
        if (FWD && tile_has_xx) {   // several extras on one node (rare): the out-of-register loop reads position-indexed copies
#pragma unroll
            for (int k = 0; k < RS_PER_LANE; k++) { const int pos = RS_PER_LANE * lane + k; sa[pos] = pm[k]; sb[pos] = pi[k]; sc[pos] = pd[k]; }
        }
        prefetch_next();   // (barrier inside) every lane is done with the raw row: refill it for the next job
        double cm[RS_PER_LANE], ci[RS_PER_LANE], dacc[RS_PER_LANE], dcur[RS_PER_LANE];
        if (FWD) {
            const double fbv = fb0.v == 0.0 ? 0.0 : fb0.v * pow2i(fb0.e - Eref);
            const double ibv = ib_cur.v == 0.0 ? 0.0 : ib_cur.v * pow2i(ib_cur.e - Eref);
            // round A: fm, fi (forward.rs:337-388)
            {
                double x0 = 0.0;
                if (tile_has_x) x0 = xt * (lp.p_MM * x0m + lp.p_IM * x0i + lp.p_DM * x0d);
                if (tile_has_xx && ext0) x0 += RS_EXTRAS0(sa, sb, sc, lp.p_MM, lp.p_IM, lp.p_DM);
#pragma unroll
                for (int k = 0; k < RS_PER_LANE; k++) {
                    const double um = k == 0 ? u0m : pm[k > 0 ? k - 1 : 0], ui = k == 0 ? u0i : pi[k > 0 ? k - 1 : 0], ud = k == 0 ? u0d : pd[k > 0 ? k - 1 : 0];
                    double acc = tr[k] * (lp.p_MM * um + lp.p_IM * ui + lp.p_DM * ud);
                    if (k == 0) acc += x0;
                    acc += fbv * sinit[32 * k + lane];
                    cm[k] = acc * (((emc >> (3 * k)) & 7) == x ? lp.p_match : lp.p_mismatch);
                    ci[k] = lp.p_random * (lp.p_MI * pm[k] + lp.p_II * pi[k] + lp.p_DI * pd[k]);
                }
            }
            if (tile_has_xx) __syncwarp();   // the position-indexed copies of the previous row have been read
            RS_STAGE(sa, cm); RS_STAGE(sb, ci);
            __syncwarp();
            // round B: fd0 (forward.rs:480-501)
            {
                double x0 = 0.0;
                if (tile_has_x) x0 = xt * (lp.p_MD * sa[xp] + lp.p_ID * sb[xp]);
                if (tile_has_xx && ext0) x0 += RS_EXTRAS0(sa, sb, sb, lp.p_MD, lp.p_ID, 0.0);
#pragma unroll
                for (int k = 0; k < RS_PER_LANE; k++) {
                    const double um = RS_UP(cm, sa, k), ui = RS_UP(ci, sb, k);
                    double acc = tr[k] * (lp.p_MD * um + lp.p_ID * ui);
                    if (k == 0) acc += x0;
                    acc += ibv * (lp.p_ID * sinit[32 * k + lane]);
                    dcur[k] = acc; dacc[k] = acc;
                }
            }
            RS_STAGE(sc, dcur);
            __syncwarp();
            // fdt x 4 (forward.rs:510-524): previous round in registers, its sources' copies in sc / sa alternately
#pragma unroll
            for (int t = 1; t < N_DEL_ROUNDS; t++) {
                double* prevbuf = (t & 1) ? sc : sa; double* curbuf = (t & 1) ? sa : sc;
                double x0 = 0.0;
                if (tile_has_x) x0 = xt * lp.p_DD * prevbuf[xp];
                if (tile_has_xx && ext0) x0 += RS_EXTRAS0(prevbuf, prevbuf, prevbuf, lp.p_DD, 0.0, 0.0);
                // in place, highest slot first: slot k reads the previous round's value of slot k - 1
#pragma unroll
                for (int k = RS_PER_LANE - 1; k >= 0; k--) {
                    double v = tr[k] * lp.p_DD * RS_UP(dcur, prevbuf, k);
                    if (k == 0) v += x0;
                    dcur[k] = v; dacc[k] += v;
                }
                if (t < N_DEL_ROUNDS - 1) RS_STAGE(curbuf, dcur);
                __syncwarp();
            }
        } else {
            // previous (= next base) row: pm := e_l(x) m''[l]  (every use of m'' is multiplied by the emission of that node)
#pragma unroll
            for (int k = 0; k < RS_PER_LANE; k++) pm[k] *= (((emc >> (3 * k)) & 7) == x ? lp.p_match : lp.p_mismatch);
            RS_STAGE(sa, pm);
            __syncwarp();
            // bd0 (backward.rs:354-377)
            {
                double x0 = 0.0;
                if (tile_has_x) x0 = xt * lp.p_DM * sa[xp];
                if (tile_has_xx && ext0) x0 += RS_EXTRAS0(sa, sa, sa, lp.p_DM, 0.0, 0.0);
#pragma unroll
                for (int k = 0; k < RS_PER_LANE; k++) {
                    double acc = tr[k] * lp.p_DM * RS_UP(pm, sa, k);
                    if (k == 0) acc += x0;
                    acc += lp.p_DI * lp.p_random * pi[k];
                    dcur[k] = acc; dacc[k] = acc;
                }
            }
            RS_STAGE(sc, dcur);
            __syncwarp();
            // bdt x 4 (backward.rs:387-404): copies alternate between sc and sb (sa keeps e m'' for bm / bi)
#pragma unroll
            for (int t = 1; t < N_DEL_ROUNDS; t++) {
                double* prevbuf = (t & 1) ? sc : sb; double* curbuf = (t & 1) ? sb : sc;
                double x0 = 0.0;
                if (tile_has_x) x0 = xt * lp.p_DD * prevbuf[xp];
                if (tile_has_xx && ext0) x0 += RS_EXTRAS0(prevbuf, prevbuf, prevbuf, lp.p_DD, 0.0, 0.0);
#pragma unroll
                for (int k = RS_PER_LANE - 1; k >= 0; k--) {
                    double v = tr[k] * lp.p_DD * RS_UP(dcur, prevbuf, k);
                    if (k == 0) v += x0;
                    dcur[k] = v; dacc[k] += v;
                }
                if (t < N_DEL_ROUNDS - 1) RS_STAGE(curbuf, dcur);
                __syncwarp();
            }
            // d of the source positions into sc, then bm / bi (backward.rs:423-483)
            RS_STAGE(sc, dacc);
            __syncwarp();
            {
                double xm = 0.0, xi_ = 0.0;
                if (tile_has_x) { xm = xt * (lp.p_MM * sa[xp] + lp.p_MD * sc[xp]); xi_ = xt * (lp.p_IM * sa[xp] + lp.p_ID * sc[xp]); }
                if (tile_has_xx && ext0) {
                    xm += RS_EXTRAS0(sa, sc, sc, lp.p_MM, lp.p_MD, 0.0);
                    xi_ += RS_EXTRAS0(sa, sc, sc, lp.p_IM, lp.p_ID, 0.0);
                }
#pragma unroll
                for (int k = 0; k < RS_PER_LANE; k++) {
                    const double um = RS_UP(pm, sa, k), ud = RS_UP(dacc, sc, k);
                    double am = tr[k] * (lp.p_MM * um + lp.p_MD * ud), ai = tr[k] * (lp.p_IM * um + lp.p_ID * ud);
                    if (k == 0) { am += xm; ai += xi_; }
                    cm[k] = am + lp.p_MI * lp.p_random * pi[k];
                    ci[k] = ai + lp.p_II * lp.p_random * pi[k];
                }
            }
            __syncwarp();
        }
        // ---- pack the owned cells, stage position-indexed, store the core coalesced; partial sums over the core
        double part = 0.0, part2 = 0.0;
#pragma unroll
        for (int k = 0; k < RS_PER_LANE; k++) {
            const int pos = RS_PER_LANE * lane + k;
            const double m = cm[k], i = ci[k], d = dacc[k];
            const double mx = fmax(m, fmax(i, d));
            int qx = 0; double scl = 0.0;
            if (mx != 0.0) { qx = ilogb_pos(mx); scl = pow2i(-qx); }
            sa[pos] = m * scl; sb[pos] = i * scl; sc[pos] = d * scl; se[pos] = mx != 0.0 ? Eref + qx : 0;
            if ((cmask >> k) & 1) {
                if (FWD) part += m + i + d;
                else { const double in = sinit[32 * k + lane]; part += (pm[k] * lp.p_MM + d * lp.p_MD) * in; part2 += (pm[k] * lp.p_IM + d * lp.p_ID) * in; }
            }
        }
        __syncwarp();
#pragma unroll
        for (int q = 0; q < RS_PER_LANE; q++) {
            const int pos = 32 * q + lane;
            if ((cmask >> (8 + q)) & 1) {
                const uint32_t g = sion[pos];
                om[g] = sa[pos]; oi[g] = sb[pos]; od[g] = sc[pos]; oe[g] = se[pos];
            }
        }
        for (int o = 16; o; o >>= 1) { part += __shfl_down_sync(0xffffffffu, part, o); if (!FWD) part2 += __shfl_down_sync(0xffffffffu, part2, o); }
        if (lane == 0) {
            if (FWD) partials[(size_t)job_idx * n_tiles + c] = xf(part, Eref);
            else { partials[((size_t)job_idx * n_tiles + c) * 2] = xf(part, Eref); partials[((size_t)job_idx * n_tiles + c) * 2 + 1] = xf(part2, Eref); }
        }
    }
}


// ------------------------------------------------------------------------------------------------ two forward rows per launch
// k_dense_fwd2: the register-stencil forward step applied twice (rows s and s + 1) on tiles with a 12-hop halo (PlanView of
// DevPlan fwd2): row s never leaves the registers, so a pair of rows costs one read and one write of the 28 B cell and one
// staging / transpose pass through shared memory instead of two.  Used when nothing needs the intermediate row in memory
// (ping-pong slabs, no per-row products): pair p reads slab (p - 1) & 1 and writes slab p & 1.
// One exponent frame serves both rows: the range check uses `span2`, which leaves room for the factors of two rows, so no
// product can underflow and the results are bit-identical to two single-row steps (power-of-two scalings are exact).
// A tile whose inputs do not fit raises *redo: the caller repeats the whole dense phase with single-row steps (rare).
template <bool FWD>
__global__ void k_dense_prep2(LinParams lp, const DJob* __restrict__ jobs, uint32_t n_jobs, uint32_t s, const uint8_t* __restrict__ bases,
                              const RowDesc* __restrict__ desc, char* __restrict__ pool, uint64_t slab_bytes, JStep* __restrict__ out) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_jobs) return;
    const DJob jb = jobs[j];
    JStep a, b;
    a.prev_ptr = 0; a.out_ptr = 0; a.fb0 = xf_zero(); a.ib_cur = xf_zero(); a.valid = 0; a.x = 0; a.pk = 0; a.base = 0;
    b = a;
    if (s < jb.n_steps) {
        const int row = FWD ? jb.first_row + (int)s : jb.first_row - (int)s;
        const uint32_t p = s >> 1;
        a.valid = 1; a.x = (int)jb.x; a.base = bases[jb.base_off + row];
        a.pk = (s == 0) ? jb.prev0_kind : PREV_SLAB;
        a.prev_ptr = (unsigned long long)(pool + (s == 0 ? jb.prev0_slab : jb.slab0 + ((p - 1) & 1)) * slab_bytes);
        a.out_ptr = (unsigned long long)(pool + (jb.slab0 + (p & 1)) * slab_bytes);
        if (FWD) {
            XF mbp, ibp;
            if (row == 0) { mbp = xf(1.0, 0); ibp = xf_zero(); }
            else { mbp = desc[jb.desc0 + row - 1].mb; ibp = desc[jb.desc0 + row - 1].ib; }
            a.ib_cur = xmul(xadd(xmul(mbp, lp.p_MI), xmul(ibp, lp.p_II)), lp.p_random);
            a.fb0 = xadd(xmul(mbp, lp.p_MM), xmul(ibp, lp.p_IM));
        }
        if (s + 1 < jb.n_steps) {
            b.valid = 1; b.x = a.x; b.base = bases[jb.base_off + (FWD ? row + 1 : row - 1)]; b.pk = PREV_SLAB; b.prev_ptr = 0; b.out_ptr = a.out_ptr;
            if (FWD) {
                // begin scalars of row s as k_dense_fwd_finish will write them: mb = 0, ib = xnorm(ib_cur)
                const XF ibs = xnorm(a.ib_cur);
                b.ib_cur = xmul(xmul(ibs, lp.p_II), lp.p_random);
                b.fb0 = xmul(ibs, lp.p_IM);
            }
        }
    }
    out[j] = a;
    out[n_jobs + j] = b;
}

#ifdef RS2_UNROLL_PASSES
#define RS2_PASS_PRAGMA _Pragma("unroll")
#else
#define RS2_PASS_PRAGMA _Pragma("unroll 1")
#endif
#define RS2_SMEM_PER_WARP (2 * DENSE_LMAX * (3 * 8 + 4) + 6 * 64 + DENSE_LMAX * (8 + 4))
#define RS2_SMEM_BYTES (WT_WARPS * RS2_SMEM_PER_WARP)

__global__ void __launch_bounds__(WT_WARPS * 32, WT_MIN_CTAS)
k_dense_fwd2(PlanView P, GraphView G, LinParams lp, const JStep* __restrict__ jstep, uint32_t n_jobs, uint32_t Np,
             XF* __restrict__ partials, uint32_t n_tiles, int span2, int* __restrict__ redo, uint32_t jpc) {
    constexpr bool FWD = true;
    extern __shared__ __align__(16) unsigned char rs_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t c = blockIdx.x * WT_WARPS + warp;
    if (c >= n_tiles) return;   // no block-wide barrier anywhere in this kernel
    double* sa = (double*)(rs_raw + (size_t)warp * RS2_SMEM_PER_WARP);
    double* sb = sa + DENSE_LMAX; double* sc = sb + DENSE_LMAX; int* se = (int*)(sc + DENSE_LMAX);
    double* rm = (double*)(se + DENSE_LMAX); double* ri = rm + DENSE_LMAX; double* rd = ri + DENSE_LMAX; int* re = (int*)(rd + DENSE_LMAX);
    JStep* sj = (JStep*)(re + DENSE_LMAX);           // [3 job slots][2 rows]
    double* sinit = (double*)(sj + 6);               // init of the owned positions, [slot][lane] (lane-private, conflict-free)
    uint32_t* sion = (uint32_t*)(sinit + DENSE_LMAX); // node of IO position 32 q + lane, [q][lane] (lane-private)
    const uint32_t tbase = c * DENSE_LMAX;
    const uint32_t g0 = P.chunk_start[c], ncore = P.chunk_start[c + 1] - g0;
    unsigned int emc = 0, cmask = 0, srcm = 0;   // as in k_dense_reg
#pragma unroll
    for (int q = 0; q < RS_PER_LANE; q++) {
        const uint32_t nd = P.rl_node[tbase + 32 * q + lane];
        sion[32 * q + lane] = nd;
        cmask |= (nd - g0 < ncore ? 1u : 0u) << (8 + q);
        if (nd != 0xffffffffu) cmask |= 1u << (16 + q);
    }
#pragma unroll
    for (int k = 0; k < RS_PER_LANE; k++) {
        const uint32_t nd = P.rl_node[tbase + RS_PER_LANE * lane + k];
        unsigned int code = 4;
        if (nd != 0xffffffffu) { unsigned char ch = G.emission[nd]; code = ch == 'n' ? 4 : ((ch >> 1) & 3); }
        emc |= code << (3 * k);
        cmask |= (nd - g0 < ncore ? 1u : 0u) << k;
        srcm |= (unsigned int)((P.rl_flag[tbase + RS_PER_LANE * lane + k] >> 2) & 1) << k;
    }
    const bool tile_plain = !__any_sync(0xffffffffu, (srcm & ((1u << (RS_PER_LANE - 1)) - 1)) != 0);
    const int pp0 = P.rl_par[tbase + RS_PER_LANE * lane];
    int xp = pp0; uint32_t xe = 0xffffffffu; double xt = 0.0; bool ext0 = false;
    if (P.rl_flag[tbase + RS_PER_LANE * lane] & 2) {
        const size_t xi = (size_t)c * (DENSE_LMAX + 1) + RS_PER_LANE * lane;
        const uint32_t a0 = P.rx_off[xi], a1 = P.rx_off[xi + 1];
        if (a1 - a0 == 1) { xp = P.rx_idx[a0]; xe = P.rx_eid[a0]; } else ext0 = true;
    }
    const bool tile_has_x = __any_sync(0xffffffffu, xe != 0xffffffffu);
    const bool tile_has_xx = __any_sync(0xffffffffu, ext0);
    double tr[RS_PER_LANE];
#pragma unroll
    for (int k = 0; k < RS_PER_LANE; k++) tr[k] = 0.0;
    int staged_x = -1;
    const double* trans = G.trans;
    const uint32_t job0 = blockIdx.y * jpc;
    const uint32_t n_here = job0 < n_jobs ? min(jpc, n_jobs - job0) : 0u;
    const size_t prow = (size_t)n_jobs * n_tiles;   // partials of the second row
    auto fetch_js = [&](uint32_t jj, int slot) {   // both rows' scalars of job jj
        if (lane < 4) cp_async16((char*)&sj[2 * slot] + 16 * lane, (const char*)&jstep[job0 + jj] + 16 * lane);
        else if (lane < 8) cp_async16((char*)&sj[2 * slot + 1] + 16 * (lane - 4), (const char*)&jstep[n_jobs + job0 + jj] + 16 * (lane - 4));
    };
    auto fetch_rows = [&](const JStep& js) {
        if (!js.valid || js.pk != PREV_SLAB) return;
        const double* gm = (const double*)js.prev_ptr; const double* gi = gm + Np; const double* gd = gi + Np; const int* ge = (const int*)(gd + Np);
#pragma unroll
        for (int q = 0; q < RS_PER_LANE; q++) {
            if ((cmask >> (16 + q)) & 1) {
                const int pos = 32 * q + lane; const uint32_t g = sion[pos];
                cp_async8(&rm[pos], &gm[g]); cp_async8(&ri[pos], &gi[g]); cp_async8(&rd[pos], &gd[g]); cp_async4(&re[pos], &ge[g]);
            }
        }
    };
    if (n_here == 0) return;
#pragma unroll
    for (int q = 0; q < RS_PER_LANE; q++) { const int pos = 32 * q + lane; rm[pos] = 0.0; ri[pos] = 0.0; rd[pos] = 0.0; re[pos] = 0; }   // pads stay zero
    fetch_js(0, 0);
    if (n_here > 1) fetch_js(1, 1);
    cp_async_wait_all();
    __syncwarp();
    fetch_rows(sj[0]);
    for (uint32_t jj = 0; jj < n_here; jj++) {
        const uint32_t job_idx = job0 + jj;
        const int slot = jj % 3;
        cp_async_wait_all();
        __syncwarp();
        // the raw row and a free JStep slot are refilled for the jobs ahead (slot (jj + 2) % 3 is not the one in use)
        auto prefetch_next = [&]() {
            __syncwarp();
            if (jj + 1 < n_here) fetch_rows(sj[2 * ((jj + 1) % 3)]);
            if (jj + 2 < n_here) fetch_js(jj + 2, (jj + 2) % 3);
        };
        const JStep* js0 = &sj[2 * slot];
        const JStep* js1 = js0 + 1;
        const int valid = js0->valid, valid2 = js1->valid, hx = js0->x, pk = js0->pk;
        if (!valid) { prefetch_next(); continue; }
        if (hx != staged_x) {
            const double* init = G.init + (size_t)hx * G.N;
            trans = G.trans + (size_t)hx * G.E;
#pragma unroll
            for (int k = 0; k < RS_PER_LANE; k++) {
                const uint32_t nd = P.rl_node[tbase + RS_PER_LANE * lane + k], ed = P.rl_eid[tbase + RS_PER_LANE * lane + k];
                sinit[32 * k + lane] = nd == 0xffffffffu ? 0.0 : init[nd];
                tr[k] = ed == 0xffffffffu ? 0.0 : trans[ed];
            }
            xt = xe == 0xffffffffu ? 0.0 : trans[xe];
            staged_x = hx;
        }
        if (pk != PREV_SLAB) {
            const double v0 = pk == PREV_B_INIT ? lp.p_end : 0.0;
#pragma unroll
            for (int q = 0; q < RS_PER_LANE; q++)
                if ((cmask >> (16 + q)) & 1) { const int pos = 32 * q + lane; rm[pos] = v0; ri[pos] = v0; rd[pos] = 0.0; re[pos] = 0; }
            __syncwarp();
        }
        double pm[RS_PER_LANE], pi[RS_PER_LANE], pd[RS_PER_LANE]; int pe[RS_PER_LANE];
        int elo = EXP_NONE_LO_, ehi = EXP_NONE_HI_;
#pragma unroll
        for (int k = 0; k < RS_PER_LANE; k++) {
            const int pos = RS_PER_LANE * lane + k;
            pm[k] = rm[pos]; pi[k] = ri[pos]; pd[k] = rd[pos]; pe[k] = re[pos];
            if (pm[k] + pi[k] + pd[k] != 0.0) { elo = pe[k] < elo ? pe[k] : elo; ehi = pe[k] > ehi ? pe[k] : ehi; }
        }
        if (lane < 4) {   // begin scalars of both rows take part in the frame
            const JStep* jq = (lane & 2) ? js1 : js0;
            const XF sv = (lane & 1) ? jq->ib_cur : jq->fb0;
            if (jq->valid && sv.v != 0.0) { int e = xexp(sv); elo = e < elo ? e : elo; ehi = e > ehi ? e : ehi; }
        }
        elo = __reduce_min_sync(0xffffffffu, elo); ehi = __reduce_max_sync(0xffffffffu, ehi);
        const unsigned long long out_ptr = js0->out_ptr;
        double* om = (double*)out_ptr; double* oi = om + Np; double* od = oi + Np; int* oe = (int*)(od + Np);
        if (ehi == EXP_NONE_HI_) {  // nothing but zeros flows into this tile: both rows are zero here
            prefetch_next();
            for (uint32_t j = lane; j < ncore; j += 32) { om[g0 + j] = 0.0; oi[g0 + j] = 0.0; od[g0 + j] = 0.0; oe[g0 + j] = 0; }
            if (lane == 0) { partials[(size_t)job_idx * n_tiles + c] = xf_zero(); if (valid2) partials[prow + (size_t)job_idx * n_tiles + c] = xf_zero(); }
            continue;
        }
        if (ehi - elo > span2) {    // exponent range too wide for one frame over two rows
            prefetch_next();
            if (lane == 0) *redo = 1;
            continue;
        }
        const int Eref = ehi;
        double u0m, u0i, u0d, x0m = 0.0, x0i = 0.0, x0d = 0.0;
        { const double s0 = pow2i(re[pp0] - Eref); u0m = rm[pp0] * s0; u0i = ri[pp0] * s0; u0d = rd[pp0] * s0; }
        if (tile_has_x) { const double sx = pow2i(re[xp] - Eref); x0m = rm[xp] * sx; x0i = ri[xp] * sx; x0d = rd[xp] * sx; }
#pragma unroll
        for (int k = 0; k < RS_PER_LANE; k++) { const double sc_ = pow2i(pe[k] - Eref); pm[k] *= sc_; pi[k] *= sc_; pd[k] *= sc_; }
        if (tile_has_xx) {
#pragma unroll
            for (int k = 0; k < RS_PER_LANE; k++) { const int pos = RS_PER_LANE * lane + k; sa[pos] = pm[k]; sb[pos] = pi[k]; sc[pos] = pd[k]; }
        }
        prefetch_next();   // (barrier inside) every lane is done with the raw row: refill it for the next job
        double cm[RS_PER_LANE], ci[RS_PER_LANE], dacc[RS_PER_LANE], dcur[RS_PER_LANE];
RS2_PASS_PRAGMA
        for (int pass = 0; pass < 2; pass++) {
            const JStep* jq = pass ? js1 : js0;
            const XF fb0 = jq->fb0, ib_cur = jq->ib_cur;
            const unsigned int x = (jq->base >> 1) & 3;   // base code of the read base
            const double fbv = fb0.v == 0.0 ? 0.0 : fb0.v * pow2i(fb0.e - Eref);
            const double ibv = ib_cur.v == 0.0 ? 0.0 : ib_cur.v * pow2i(ib_cur.e - Eref);
            // round A: fm, fi (forward.rs:337-388)
            {
                double x0 = 0.0;
                if (tile_has_x) x0 = xt * (lp.p_MM * x0m + lp.p_IM * x0i + lp.p_DM * x0d);
                if (tile_has_xx && ext0) x0 += RS_EXTRAS0(sa, sb, sc, lp.p_MM, lp.p_IM, lp.p_DM);
#pragma unroll
                for (int k = 0; k < RS_PER_LANE; k++) {
                    const double um = k == 0 ? u0m : pm[k > 0 ? k - 1 : 0], ui = k == 0 ? u0i : pi[k > 0 ? k - 1 : 0], ud = k == 0 ? u0d : pd[k > 0 ? k - 1 : 0];
                    double acc = tr[k] * (lp.p_MM * um + lp.p_IM * ui + lp.p_DM * ud);
                    if (k == 0) acc += x0;
                    acc += fbv * sinit[32 * k + lane];
                    cm[k] = acc * (((emc >> (3 * k)) & 7) == x ? lp.p_match : lp.p_mismatch);
                    ci[k] = lp.p_random * (lp.p_MI * pm[k] + lp.p_II * pi[k] + lp.p_DI * pd[k]);
                }
            }
            if (tile_has_xx) __syncwarp();   // the position-indexed copies of the previous row have been read
            RS_STAGE(sa, cm); RS_STAGE(sb, ci);
            __syncwarp();
            // round B: fd0 (forward.rs:480-501)
            {
                double x0 = 0.0;
                if (tile_has_x) x0 = xt * (lp.p_MD * sa[xp] + lp.p_ID * sb[xp]);
                if (tile_has_xx && ext0) x0 += RS_EXTRAS0(sa, sb, sb, lp.p_MD, lp.p_ID, 0.0);
#pragma unroll
                for (int k = 0; k < RS_PER_LANE; k++) {
                    const double um = RS_UP(cm, sa, k), ui = RS_UP(ci, sb, k);
                    double acc = tr[k] * (lp.p_MD * um + lp.p_ID * ui);
                    if (k == 0) acc += x0;
                    acc += ibv * (lp.p_ID * sinit[32 * k + lane]);
                    dcur[k] = acc; dacc[k] = acc;
                }
            }
            RS_STAGE(sc, dcur);
            __syncwarp();
            // fdt x 4 (forward.rs:510-524)
#pragma unroll
            for (int t = 1; t < N_DEL_ROUNDS; t++) {
                double* prevbuf = (t & 1) ? sc : sa; double* curbuf = (t & 1) ? sa : sc;
                double x0 = 0.0;
                if (tile_has_x) x0 = xt * lp.p_DD * prevbuf[xp];
                if (tile_has_xx && ext0) x0 += RS_EXTRAS0(prevbuf, prevbuf, prevbuf, lp.p_DD, 0.0, 0.0);
#pragma unroll
                for (int k = RS_PER_LANE - 1; k >= 0; k--) {
                    double v = tr[k] * lp.p_DD * RS_UP(dcur, prevbuf, k);
                    if (k == 0) v += x0;
                    dcur[k] = v; dacc[k] += v;
                }
                if (t < N_DEL_ROUNDS - 1) RS_STAGE(curbuf, dcur);
                __syncwarp();
            }
            // fe partial of this row over the core
            double part = 0.0;
#pragma unroll
            for (int k = 0; k < RS_PER_LANE; k++) if ((cmask >> k) & 1) part += cm[k] + ci[k] + dacc[k];
            for (int o = 16; o; o >>= 1) part += __shfl_down_sync(0xffffffffu, part, o);
            if (lane == 0) partials[(pass ? prow : 0) + (size_t)job_idx * n_tiles + c] = xf(part, Eref);
            if (pass == 1 || !valid2) break;
            // ---- row s becomes the previous row: publish what other lanes read, take the slot-0 neighbours, go again
            if (tile_has_xx) {
#pragma unroll
                for (int k = 0; k < RS_PER_LANE; k++) { const int pos = RS_PER_LANE * lane + k; sa[pos] = cm[k]; sb[pos] = ci[k]; sc[pos] = dacc[k]; }
            } else { RS_STAGE(sa, cm); RS_STAGE(sb, ci); RS_STAGE(sc, dacc); }
            __syncwarp();
            u0m = sa[pp0]; u0i = sb[pp0]; u0d = sc[pp0];
            if (tile_has_x) { x0m = sa[xp]; x0i = sb[xp]; x0d = sc[xp]; }
#pragma unroll
            for (int k = 0; k < RS_PER_LANE; k++) { pm[k] = cm[k]; pi[k] = ci[k]; pd[k] = dacc[k]; }
            if (!tile_has_xx) __syncwarp();   // (with extras the copies stay until round A has read them: barrier there)
        }
        // ---- pack the owned cells, stage position-indexed, store the core coalesced
#pragma unroll
        for (int k = 0; k < RS_PER_LANE; k++) {
            const int pos = RS_PER_LANE * lane + k;
            const double m = cm[k], i = ci[k], d = dacc[k];
            // exponent of the largest state: the values are non-negative, so the largest high word carries it
            const int hx = max(__double2hiint(m), max(__double2hiint(i), __double2hiint(d)));
            const bool nz = m + i + d != 0.0;
            int qx = 0; double scl = 0.0;
            if (nz) { qx = (hx >> 20) - 1023; scl = pow2i(-qx); }
            sa[pos] = m * scl; sb[pos] = i * scl; sc[pos] = d * scl; se[pos] = nz ? Eref + qx : 0;
        }
        __syncwarp();
#pragma unroll
        for (int q = 0; q < RS_PER_LANE; q++) {
            const int pos = 32 * q + lane;
            if ((cmask >> (8 + q)) & 1) {
                const uint32_t g = sion[pos];
                om[g] = sa[pos]; oi[g] = sb[pos]; od[g] = sc[pos]; oe[g] = se[pos];
            }
        }
        __syncwarp();   // (the next job's round A stages into sa / sb)
    }
}


// k_dense_bwd2: two backward rows (r, r - 1) per launch, the mirror of k_dense_fwd2 on the tiles of DevPlan bwd2.
__global__ void __launch_bounds__(WT_WARPS * 32, WT_MIN_CTAS)
k_dense_bwd2(PlanView P, GraphView G, LinParams lp, const JStep* __restrict__ jstep, uint32_t n_jobs, uint32_t Np,
             XF* __restrict__ partials, uint32_t n_tiles, int span2, int* __restrict__ redo, uint32_t jpc) {
    extern __shared__ __align__(16) unsigned char rs_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t c = blockIdx.x * WT_WARPS + warp;
    if (c >= n_tiles) return;
    double* sa = (double*)(rs_raw + (size_t)warp * RS2_SMEM_PER_WARP);
    double* sb = sa + DENSE_LMAX; double* sc = sb + DENSE_LMAX; int* se = (int*)(sc + DENSE_LMAX);
    double* rm = (double*)(se + DENSE_LMAX); double* ri = rm + DENSE_LMAX; double* rd = ri + DENSE_LMAX; int* re = (int*)(rd + DENSE_LMAX);
    JStep* sj = (JStep*)(re + DENSE_LMAX);           // [3 job slots][2 rows]
    double* sinit = (double*)(sj + 6);
    uint32_t* sion = (uint32_t*)(sinit + DENSE_LMAX);
    (void)rd;
    const uint32_t tbase = c * DENSE_LMAX;
    const uint32_t g0 = P.chunk_start[c], ncore = P.chunk_start[c + 1] - g0;
    unsigned int emc = 0, cmask = 0, srcm = 0;
#pragma unroll
    for (int q = 0; q < RS_PER_LANE; q++) {
        const uint32_t nd = P.rl_node[tbase + 32 * q + lane];
        sion[32 * q + lane] = nd;
        cmask |= (nd - g0 < ncore ? 1u : 0u) << (8 + q);
        if (nd != 0xffffffffu) cmask |= 1u << (16 + q);
    }
#pragma unroll
    for (int k = 0; k < RS_PER_LANE; k++) {
        const uint32_t nd = P.rl_node[tbase + RS_PER_LANE * lane + k];
        unsigned int code = 4;
        if (nd != 0xffffffffu) { unsigned char ch = G.emission[nd]; code = ch == 'n' ? 4 : ((ch >> 1) & 3); }
        emc |= code << (3 * k);
        cmask |= (nd - g0 < ncore ? 1u : 0u) << k;
        srcm |= (unsigned int)((P.rl_flag[tbase + RS_PER_LANE * lane + k] >> 2) & 1) << k;
    }
    const bool tile_plain = !__any_sync(0xffffffffu, (srcm & ((1u << (RS_PER_LANE - 1)) - 1)) != 0);
    const int pp0 = P.rl_par[tbase + RS_PER_LANE * lane];
    int xp = pp0; uint32_t xe = 0xffffffffu; double xt = 0.0; bool ext0 = false;
    if (P.rl_flag[tbase + RS_PER_LANE * lane] & 2) {
        const size_t xi = (size_t)c * (DENSE_LMAX + 1) + RS_PER_LANE * lane;
        const uint32_t a0 = P.rx_off[xi], a1 = P.rx_off[xi + 1];
        if (a1 - a0 == 1) { xp = P.rx_idx[a0]; xe = P.rx_eid[a0]; } else ext0 = true;
    }
    const bool tile_has_x = __any_sync(0xffffffffu, xe != 0xffffffffu);
    const bool tile_has_xx = __any_sync(0xffffffffu, ext0);
    double tr[RS_PER_LANE];
#pragma unroll
    for (int k = 0; k < RS_PER_LANE; k++) tr[k] = 0.0;
    int staged_x = -1;
    const double* trans = G.trans;
    const uint32_t job0 = blockIdx.y * jpc;
    const uint32_t n_here = job0 < n_jobs ? min(jpc, n_jobs - job0) : 0u;
    const size_t prow = (size_t)n_jobs * n_tiles * 2;   // partials of the second row
    auto fetch_js = [&](uint32_t jj, int slot) {
        if (lane < 4) cp_async16((char*)&sj[2 * slot] + 16 * lane, (const char*)&jstep[job0 + jj] + 16 * lane);
        else if (lane < 8) cp_async16((char*)&sj[2 * slot + 1] + 16 * (lane - 4), (const char*)&jstep[n_jobs + job0 + jj] + 16 * (lane - 4));
    };
    auto fetch_rows = [&](const JStep& js) {
        if (!js.valid || js.pk != PREV_SLAB) return;
        const double* gm = (const double*)js.prev_ptr; const double* gi = gm + Np; const int* ge = (const int*)(gi + 2 * (size_t)Np);
#pragma unroll
        for (int q = 0; q < RS_PER_LANE; q++) {
            if ((cmask >> (16 + q)) & 1) {
                const int pos = 32 * q + lane; const uint32_t g = sion[pos];
                cp_async8(&rm[pos], &gm[g]); cp_async8(&ri[pos], &gi[g]); cp_async4(&re[pos], &ge[g]);
            }
        }
    };
    if (n_here == 0) return;
#pragma unroll
    for (int q = 0; q < RS_PER_LANE; q++) { const int pos = 32 * q + lane; rm[pos] = 0.0; ri[pos] = 0.0; re[pos] = 0; }
    fetch_js(0, 0);
    if (n_here > 1) fetch_js(1, 1);
    cp_async_wait_all();
    __syncwarp();
    fetch_rows(sj[0]);
    for (uint32_t jj = 0; jj < n_here; jj++) {
        const uint32_t job_idx = job0 + jj;
        const int slot = jj % 3;
        cp_async_wait_all();
        __syncwarp();
        auto prefetch_next = [&]() {
            __syncwarp();
            if (jj + 1 < n_here) fetch_rows(sj[2 * ((jj + 1) % 3)]);
            if (jj + 2 < n_here) fetch_js(jj + 2, (jj + 2) % 3);
        };
        const JStep* js0 = &sj[2 * slot];
        const JStep* js1 = js0 + 1;
        const int valid = js0->valid, valid2 = js1->valid, hx = js0->x, pk = js0->pk;
        if (!valid) { prefetch_next(); continue; }
        if (hx != staged_x) {
            const double* init = G.init + (size_t)hx * G.N;
            trans = G.trans + (size_t)hx * G.E;
#pragma unroll
            for (int k = 0; k < RS_PER_LANE; k++) {
                const uint32_t nd = P.rl_node[tbase + RS_PER_LANE * lane + k], ed = P.rl_eid[tbase + RS_PER_LANE * lane + k];
                sinit[32 * k + lane] = nd == 0xffffffffu ? 0.0 : init[nd];
                tr[k] = ed == 0xffffffffu ? 0.0 : trans[ed];
            }
            xt = xe == 0xffffffffu ? 0.0 : trans[xe];
            staged_x = hx;
        }
        if (pk != PREV_SLAB) {
            const double v0 = pk == PREV_B_INIT ? lp.p_end : 0.0;
#pragma unroll
            for (int q = 0; q < RS_PER_LANE; q++)
                if ((cmask >> (16 + q)) & 1) { const int pos = 32 * q + lane; rm[pos] = v0; ri[pos] = v0; re[pos] = 0; }
            __syncwarp();
        }
        double pm[RS_PER_LANE], pi[RS_PER_LANE]; int pe[RS_PER_LANE];
        int elo = EXP_NONE_LO_, ehi = EXP_NONE_HI_;
#pragma unroll
        for (int k = 0; k < RS_PER_LANE; k++) {
            const int pos = RS_PER_LANE * lane + k;
            pm[k] = rm[pos]; pi[k] = ri[pos]; pe[k] = re[pos];
            if (pm[k] + pi[k] != 0.0) { elo = pe[k] < elo ? pe[k] : elo; ehi = pe[k] > ehi ? pe[k] : ehi; }
        }
        elo = __reduce_min_sync(0xffffffffu, elo); ehi = __reduce_max_sync(0xffffffffu, ehi);
        const unsigned long long out_ptr = js0->out_ptr;
        double* om = (double*)out_ptr; double* oi = om + Np; double* od = oi + Np; int* oe = (int*)(od + Np);
        if (ehi == EXP_NONE_HI_) {  // nothing but zeros flows into this tile: both rows are zero here
            prefetch_next();
            for (uint32_t j = lane; j < ncore; j += 32) { om[g0 + j] = 0.0; oi[g0 + j] = 0.0; od[g0 + j] = 0.0; oe[g0 + j] = 0; }
            if (lane < 2) { partials[((size_t)job_idx * n_tiles + c) * 2 + lane] = xf_zero(); if (valid2) partials[prow + ((size_t)job_idx * n_tiles + c) * 2 + lane] = xf_zero(); }
            continue;
        }
        if (ehi - elo > span2) {
            prefetch_next();
            if (lane == 0) *redo = 1;
            continue;
        }
        const int Eref = ehi;
#pragma unroll
        for (int k = 0; k < RS_PER_LANE; k++) { const double sc_ = pow2i(pe[k] - Eref); pm[k] *= sc_; pi[k] *= sc_; }
        prefetch_next();
        double cm[RS_PER_LANE], ci[RS_PER_LANE], dacc[RS_PER_LANE], dcur[RS_PER_LANE];
RS2_PASS_PRAGMA
        for (int pass = 0; pass < 2; pass++) {
            const unsigned int x = ((pass ? js1 : js0)->base >> 1) & 3;
            // next-base row: pm := e_l(x) m''[l]  (every use of m'' is multiplied by the emission of that node)
#pragma unroll
            for (int k = 0; k < RS_PER_LANE; k++) pm[k] *= (((emc >> (3 * k)) & 7) == x ? lp.p_match : lp.p_mismatch);
            RS_STAGE(sa, pm);
            __syncwarp();
            // bd0 (backward.rs:354-377)
            {
                double x0 = 0.0;
                if (tile_has_x) x0 = xt * lp.p_DM * sa[xp];
                if (tile_has_xx && ext0) x0 += RS_EXTRAS0(sa, sa, sa, lp.p_DM, 0.0, 0.0);
#pragma unroll
                for (int k = 0; k < RS_PER_LANE; k++) {
                    double acc = tr[k] * lp.p_DM * RS_UP(pm, sa, k);
                    if (k == 0) acc += x0;
                    acc += lp.p_DI * lp.p_random * pi[k];
                    dcur[k] = acc; dacc[k] = acc;
                }
            }
            RS_STAGE(sc, dcur);
            __syncwarp();
            // bdt x 4 (backward.rs:387-404): copies alternate between sc and sb (sa keeps e m'' for bm / bi)
#pragma unroll
            for (int t = 1; t < N_DEL_ROUNDS; t++) {
                double* prevbuf = (t & 1) ? sc : sb; double* curbuf = (t & 1) ? sb : sc;
                double x0 = 0.0;
                if (tile_has_x) x0 = xt * lp.p_DD * prevbuf[xp];
                if (tile_has_xx && ext0) x0 += RS_EXTRAS0(prevbuf, prevbuf, prevbuf, lp.p_DD, 0.0, 0.0);
#pragma unroll
                for (int k = RS_PER_LANE - 1; k >= 0; k--) {
                    double v = tr[k] * lp.p_DD * RS_UP(dcur, prevbuf, k);
                    if (k == 0) v += x0;
                    dcur[k] = v; dacc[k] += v;
                }
                if (t < N_DEL_ROUNDS - 1) RS_STAGE(curbuf, dcur);
                __syncwarp();
            }
            // d of the source positions into sc, then bm / bi (backward.rs:423-483)
            RS_STAGE(sc, dacc);
            __syncwarp();
            {
                double xm = 0.0, xi_ = 0.0;
                if (tile_has_x) { xm = xt * (lp.p_MM * sa[xp] + lp.p_MD * sc[xp]); xi_ = xt * (lp.p_IM * sa[xp] + lp.p_ID * sc[xp]); }
                if (tile_has_xx && ext0) {
                    xm += RS_EXTRAS0(sa, sc, sc, lp.p_MM, lp.p_MD, 0.0);
                    xi_ += RS_EXTRAS0(sa, sc, sc, lp.p_IM, lp.p_ID, 0.0);
                }
#pragma unroll
                for (int k = 0; k < RS_PER_LANE; k++) {
                    const double um = RS_UP(pm, sa, k), ud = RS_UP(dacc, sc, k);
                    double am = tr[k] * (lp.p_MM * um + lp.p_MD * ud), ai = tr[k] * (lp.p_IM * um + lp.p_ID * ud);
                    if (k == 0) { am += xm; ai += xi_; }
                    cm[k] = am + lp.p_MI * lp.p_random * pi[k];
                    ci[k] = ai + lp.p_II * lp.p_random * pi[k];
                }
            }
            // begin sums of this row over the core: init_l (p_XM e_l(x) m''[l] + p_XD d[l])  (backward.rs:499-555)
            double part = 0.0, part2 = 0.0;
#pragma unroll
            for (int k = 0; k < RS_PER_LANE; k++)
                if ((cmask >> k) & 1) { const double in = sinit[32 * k + lane]; part += (pm[k] * lp.p_MM + dacc[k] * lp.p_MD) * in; part2 += (pm[k] * lp.p_IM + dacc[k] * lp.p_ID) * in; }
            for (int o = 16; o; o >>= 1) { part += __shfl_down_sync(0xffffffffu, part, o); part2 += __shfl_down_sync(0xffffffffu, part2, o); }
            if (lane == 0) {
                XF* pp = partials + (pass ? prow : 0) + ((size_t)job_idx * n_tiles + c) * 2;
                pp[0] = xf(part, Eref); pp[1] = xf(part2, Eref);
            }
            __syncwarp();   // sa / sc have been read by every lane
            if (pass == 1 || !valid2) break;
#pragma unroll
            for (int k = 0; k < RS_PER_LANE; k++) { pm[k] = cm[k]; pi[k] = ci[k]; }
        }
        // ---- pack the owned cells, stage position-indexed, store the core coalesced
#pragma unroll
        for (int k = 0; k < RS_PER_LANE; k++) {
            const int pos = RS_PER_LANE * lane + k;
            const double m = cm[k], i = ci[k], d = dacc[k];
            // exponent of the largest state: the values are non-negative, so the largest high word carries it
            const int hx = max(__double2hiint(m), max(__double2hiint(i), __double2hiint(d)));
            const bool nz = m + i + d != 0.0;
            int qx = 0; double scl = 0.0;
            if (nz) { qx = (hx >> 20) - 1023; scl = pow2i(-qx); }
            sa[pos] = m * scl; sb[pos] = i * scl; sc[pos] = d * scl; se[pos] = nz ? Eref + qx : 0;
        }
        __syncwarp();
#pragma unroll
        for (int q = 0; q < RS_PER_LANE; q++) {
            const int pos = 32 * q + lane;
            if ((cmask >> (8 + q)) & 1) {
                const uint32_t g = sion[pos];
                om[g] = sa[pos]; oi[g] = sb[pos]; od[g] = sc[pos]; oe[g] = se[pos];
            }
        }
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------------ top-k selection
// Key of a cell: (T, mantissa, ~original id) with merged value v = (m+i+d) * 2^ex = mant * 2^T, mant in [1,2).
// Exact ordering of the linear values; ties -> lower ORIGINAL node index first (the dense SparseVec iterates in
// node-index order; UNPINNED in the reference, see oracle header).  Multi-level radix select over the digits
// [T window | 5 mantissa digits | 3 id digits]; as soon as the boundary bin holds <= SELECT_CAP cells the rest is
// resolved in shared memory.
struct SKey { int T; unsigned long long mant; uint32_t inv_id; };
__device__ __forceinline__ bool skey_gt(const SKey& a, const SKey& b) {
    if (a.T != b.T) return a.T > b.T;
    if (a.mant != b.mant) return a.mant > b.mant;
    return a.inv_id > b.inv_id;
}
__device__ __forceinline__ SKey cell_key(double m, double i, double d, int ex, uint32_t orig) {
    double v = m + i + d;
    SKey k;
    k.inv_id = ~orig;
    if (v == 0.0) { k.T = XF_ZERO_E; k.mant = 0; return k; }
    long long b = __double_as_longlong(v);
    k.T = ex + (int)((b >> 52) & 0x7ff) - 1023;
    k.mant = (unsigned long long)b & 0xfffffffffffffull;
    return k;
}
// digit `lvl` (0..8) of a key relative to window top Ttop: 11 bits each, larger = better
#define SEL_BINS 2048
__device__ __forceinline__ int skey_digit(const SKey& k, int lvl, int Ttop) {
    if (lvl == 0) { long long dlt = (long long)Ttop - k.T; return dlt >= SEL_BINS - 1 ? 0 : (int)(SEL_BINS - 1 - dlt); }  // bin 0 = everything far below
    if (lvl <= 4) return (int)((k.mant >> (52 - 11 * lvl)) & 0x7ff);
    if (lvl == 5) return (int)(k.mant & 0xff);
    if (lvl == 6) return (int)((k.inv_id >> 21) & 0x7ff);
    if (lvl == 7) return (int)((k.inv_id >> 10) & 0x7ff);
    return (int)(k.inv_id & 0x3ff);
}

// Pre-filter of the selection on large rows.  One streaming pass (k_select_tilemax) leaves the largest merged value of every run of
// SEL_TILE consecutive cells ; the selection kernel then finds a threshold tau that at least k of those tile maxima reach -- every
// one of them is a cell, so the k-th largest cell is >= tau -- and looks only into the tiles whose maximum reaches it (about k of
// N / SEL_TILE) instead of sweeping the row three or more times.  Order-preserving 64-bit image of (T, mant) relative to the row
// maximum Ttop: zero and everything more than 2046 binades below map to 0 (if tau ends up 0 the full algorithm takes over).
#define SEL_TILE 512
struct TileKey { unsigned long long mant; int T; int pad; };
__device__ __forceinline__ unsigned long long skey_pack(int T, unsigned long long mant, int Ttop) {
    if (T == XF_ZERO_E) return 0ull;
    const long long rel = (long long)Ttop - T;
    if (rel < 0 || rel > 2046) return 0ull;
    return ((unsigned long long)(2047 - rel) << 52) | mant;
}
__global__ void __launch_bounds__(256)
k_select_tilemax(const SelectReq* __restrict__ reqs, const int* __restrict__ active, const char* __restrict__ pool, uint64_t slab_bytes,
                 uint32_t Np, uint32_t N, uint32_t n_tiles, TileKey* __restrict__ tiles) {
    const SelectReq rq = reqs[blockIdx.y];
    if (rq.active_idx >= 0 && !active[rq.active_idx]) return;
    const int lane = threadIdx.x & 31;
    const uint32_t t = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (t >= n_tiles) return;
    const char* sl = pool + rq.slab * slab_bytes;
    const double* gm = (const double*)sl; const double* gi = gm + Np; const double* gd = gi + Np; const int* ge = (const int*)(gd + Np);
    int T = XF_ZERO_E; unsigned long long mant = 0;
    const uint32_t g0 = t * SEL_TILE;
#pragma unroll 4
    for (int j = 0; j < SEL_TILE / 32; j++) {
        const uint32_t g = g0 + 32 * j + lane;
        if (g < N) {
            const SKey k = cell_key(gm[g], gi[g], gd[g], ge[g], 0);
            if (k.T > T || (k.T == T && k.mant > mant)) { T = k.T; mant = k.mant; }
        }
    }
    const int Tm = __reduce_max_sync(0xffffffffu, T);
    const unsigned hi = T == Tm ? (unsigned)(mant >> 32) : 0u;
    const unsigned hm = __reduce_max_sync(0xffffffffu, hi);
    const unsigned lo = (T == Tm && hi == hm) ? (unsigned)mant : 0u;
    const unsigned lm = __reduce_max_sync(0xffffffffu, lo);
    if (lane == 0) { TileKey k; k.mant = ((unsigned long long)hm << 32) | lm; k.T = Tm; k.pad = 0; tiles[(size_t)blockIdx.y * n_tiles + t] = k; }
}

#define SELECT_SMEM_BYTES (SELECT_CAP * (8 + 4 + 4 + 4))
__global__ void __launch_bounds__(SELECT_THREADS, 1)
k_dense_select(GraphView G, const SelectReq* __restrict__ reqs, const int* __restrict__ active, const char* __restrict__ pool,
               uint64_t slab_bytes, uint32_t Np, uint32_t* __restrict__ out_ids, uint32_t* __restrict__ out_cnt,
               const TileKey* __restrict__ tiles, uint32_t n_tiles) {
    const SelectReq rq = reqs[blockIdx.x];
    if (rq.active_idx >= 0 && !active[rq.active_idx]) return;
    const char* sl = pool + rq.slab * slab_bytes;
    const double* gm = (const double*)sl; const double* gi = gm + Np; const double* gd = gi + Np; const int* ge = (const int*)(gd + Np);
    const uint32_t N = G.N;
    const int tid = threadIdx.x;
    extern __shared__ __align__(16) unsigned char sel_smem[];
    unsigned long long* c_mant = (unsigned long long*)sel_smem;
    int* c_T = (int*)(c_mant + SELECT_CAP);
    uint32_t* c_inv = (uint32_t*)(c_T + SELECT_CAP);
    uint32_t* c_node = c_inv + SELECT_CAP;
    __shared__ unsigned int hist[SEL_BINS];
    __shared__ int sh_T[SELECT_THREADS / 32];
    __shared__ int prefix[9];  // chosen digit per level
    __shared__ int s_Ttop;
    __shared__ unsigned int s_need, s_ncand, s_done, s_shift;
    __shared__ double s_L0;
    const uint32_t K = rq.k < N ? rq.k : N;
    bool have = false;   // the candidate arrays hold a superset of the top K (tile pre-filter)
    if (tiles) {
        const TileKey* tk = tiles + (size_t)blockIdx.x * n_tiles;
        __shared__ unsigned long long s_tau;
        int tmax = XF_ZERO_E;
        for (uint32_t t = tid; t < n_tiles; t += SELECT_THREADS) { const int T = tk[t].T; tmax = T > tmax ? T : tmax; }
        tmax = __reduce_max_sync(0xffffffffu, tmax);
        if ((tid & 31) == 0) sh_T[tid >> 5] = tmax;
        for (int b = tid; b < SEL_BINS; b += SELECT_THREADS) hist[b] = 0;
        __syncthreads();
        tmax = sh_T[0];
        for (int w = 1; w < SELECT_THREADS / 32; w++) tmax = sh_T[w] > tmax ? sh_T[w] : tmax;
        const int Tt = tmax;
        // two radix levels over the tile maxima: exponent digit, then the top 11 mantissa bits
        for (uint32_t t = tid; t < n_tiles; t += SELECT_THREADS) atomicAdd(&hist[(int)(skey_pack(tk[t].T, tk[t].mant, Tt) >> 52)], 1u);
        __syncthreads();
        if (tid == 0) {
            unsigned int cum = 0;
            int b = SEL_BINS - 1;
            for (; b > 0; b--) { if (cum + hist[b] >= K) break; cum += hist[b]; }
            prefix[0] = b; s_need = K - cum;
        }
        __syncthreads();
        const int b0 = prefix[0];
        for (int b = tid; b < SEL_BINS; b += SELECT_THREADS) hist[b] = 0;
        __syncthreads();
        if (b0 > 0)
            for (uint32_t t = tid; t < n_tiles; t += SELECT_THREADS) {
                const unsigned long long pk = skey_pack(tk[t].T, tk[t].mant, Tt);
                if ((int)(pk >> 52) == b0) atomicAdd(&hist[(int)((pk >> 41) & 0x7ff)], 1u);
            }
        __syncthreads();
        if (tid == 0) {
            unsigned int need = s_need, cum = 0;
            int b = SEL_BINS - 1;
            for (; b > 0; b--) { if (cum + hist[b] >= need) break; cum += hist[b]; }
            s_tau = b0 > 0 ? (((unsigned long long)b0 << 52) | ((unsigned long long)b << 41)) : 0ull;
            s_ncand = 0;
        }
        __syncthreads();
        const unsigned long long tau = s_tau;
        if (tau) {
            for (uint32_t tb = (uint32_t)(tid >> 5) * 32; tb < n_tiles; tb += SELECT_THREADS) {   // a warp takes 32 tiles at a time
                const uint32_t t = tb + (tid & 31);
                unsigned int mask = __ballot_sync(0xffffffffu, t < n_tiles && skey_pack(tk[t < n_tiles ? t : 0].T, tk[t < n_tiles ? t : 0].mant, Tt) >= tau);
                while (mask) {
                    const uint32_t g0 = (tb + (uint32_t)(__ffs(mask) - 1)) * SEL_TILE;
                    mask &= mask - 1;
                    for (int j = 0; j < SEL_TILE / 32; j++) {
                        const uint32_t g = g0 + 32 * j + (tid & 31);
                        if (g >= N) break;
                        const SKey k = cell_key(gm[g], gi[g], gd[g], ge[g], G.orig_of[g]);
                        if (skey_pack(k.T, k.mant, Tt) >= tau) {
                            const unsigned int slot = atomicAdd(&s_ncand, 1u);
                            if (slot < SELECT_CAP) { c_mant[slot] = k.mant; c_T[slot] = k.T; c_inv[slot] = k.inv_id; c_node[slot] = g; }
                        }
                    }
                }
            }
            __syncthreads();
            have = s_ncand >= K && s_ncand <= SELECT_CAP;   // (at least K by construction ; more than fit: the full algorithm below)
        }
        __syncthreads();
    }
    int nlev = 0;
    if (!have) {
    if (tid == 0) { s_need = K; s_ncand = 0; s_Ttop = 0x7fffffff; }
    __syncthreads();
    // ---- level 0: exponent window [Ttop-2046, Ttop]; shift the window down while fewer than `need` cells are in it
    for (;;) {
        const int Tcap = s_Ttop;  // only cells with T <= Tcap (strictly below the previous window) remain
        int tmax = XF_ZERO_E;
        for (uint32_t g = tid; g < N; g += SELECT_THREADS) {
            SKey k = cell_key(gm[g], gi[g], gd[g], ge[g], 0);
            if (k.T <= Tcap || Tcap == 0x7fffffff) tmax = k.T > tmax ? k.T : tmax;
        }
        for (int o = 16; o; o >>= 1) { int t = __shfl_down_sync(0xffffffffu, tmax, o); tmax = t > tmax ? t : tmax; }
        if ((tid & 31) == 0) sh_T[tid >> 5] = tmax;
        for (int b = tid; b < SEL_BINS; b += SELECT_THREADS) hist[b] = 0;
        __syncthreads();
        if (tid == 0) {
            int t = XF_ZERO_E;
            for (int w = 0; w < SELECT_THREADS / 32; w++) t = sh_T[w] > t ? sh_T[w] : t;
            s_Ttop = t;
        }
        __syncthreads();
        const int Ttop = s_Ttop;
        for (uint32_t g = tid; g < N; g += SELECT_THREADS) {
            SKey k = cell_key(gm[g], gi[g], gd[g], ge[g], 0);
            if (k.T <= Ttop) atomicAdd(&hist[skey_digit(k, 0, Ttop)], 1u);
        }
        __syncthreads();
        if (tid == 0) {
            unsigned int need = s_need, cum = 0;
            int b = SEL_BINS - 1;
            for (; b > 0; b--) { if (cum + hist[b] >= need) break; cum += hist[b]; }
            s_shift = 0;
            if (b == 0 && Ttop != XF_ZERO_E) {  // not enough cells in this window: all of bins >= 1 win, look further down
                s_need = need - cum;
                s_Ttop = Ttop - (SEL_BINS - 1);
                s_shift = 1;
            } else {
                prefix[0] = b;
                s_need = need - cum;
                s_done = (hist[b] + (K - s_need) <= SELECT_CAP) ? 1u : 0u;
            }
        }
        __syncthreads();
        if (!s_shift) break;
    }
    const int Ttop = s_Ttop;
    nlev = 1;
    // ---- levels 1..8: mantissa digits, then inverted original id digits
    for (int lvl = 1; lvl < 9 && !s_done; lvl++) {
        for (int b = tid; b < SEL_BINS; b += SELECT_THREADS) hist[b] = 0;
        __syncthreads();
        for (uint32_t g = tid; g < N; g += SELECT_THREADS) {
            SKey k = cell_key(gm[g], gi[g], gd[g], ge[g], G.orig_of[g]);
            bool match = k.T <= Ttop;
            for (int l2 = 0; l2 < lvl; l2++) match = match && (skey_digit(k, l2, Ttop) == prefix[l2]);
            if (match) atomicAdd(&hist[skey_digit(k, lvl, Ttop)], 1u);
        }
        __syncthreads();
        if (tid == 0) {
            unsigned int need = s_need, cum = 0;
            int b = SEL_BINS - 1;
            for (; b > 0; b--) { if (cum + hist[b] >= need) break; cum += hist[b]; }
            prefix[lvl] = b;
            s_need = need - cum;
            s_done = (hist[b] + (K - s_need) <= SELECT_CAP) ? 1u : 0u;
        }
        nlev = lvl + 1;
        __syncthreads();
    }
    // ---- collect every cell whose key is >= the boundary prefix (sure winners + boundary bin)
    for (uint32_t g = tid; g < N; g += SELECT_THREADS) {
        SKey k = cell_key(gm[g], gi[g], gd[g], ge[g], G.orig_of[g]);
        bool take = true;
        if (k.T <= Ttop) {
            for (int l2 = 0; l2 < nlev; l2++) {
                int dg = skey_digit(k, l2, Ttop);
                if (dg > prefix[l2]) { take = true; break; }
                if (dg < prefix[l2]) { take = false; break; }
            }
        }
        if (take) {
            unsigned int slot = atomicAdd(&s_ncand, 1u);
            if (slot < SELECT_CAP) { c_mant[slot] = k.mant; c_T[slot] = k.T; c_inv[slot] = k.inv_id; c_node[slot] = g; }
        }
    }
    __syncthreads();
    }   // !have
    const unsigned int nc = s_ncand < SELECT_CAP ? s_ncand : SELECT_CAP;
    // ---- rank by counting (nc <= SELECT_CAP), emit in descending order
    for (unsigned int a = tid; a < nc; a += SELECT_THREADS) {
        SKey ka; ka.T = c_T[a]; ka.mant = c_mant[a]; ka.inv_id = c_inv[a];
        unsigned int rank = 0;
        for (unsigned int b = 0; b < nc; b++) {
            SKey kb; kb.T = c_T[b]; kb.mant = c_mant[b]; kb.inv_id = c_inv[b];
            rank += skey_gt(kb, ka) ? 1u : 0u;
        }
        if (rank < K) out_ids[(size_t)rq.out * MAX_ACTIVE + rank] = c_node[a];
        if (rank == 0) {
            uint32_t g = c_node[a];
            s_L0 = xlog(xf(gm[g] + gi[g] + gd[g], ge[g]));
        }
    }
    __syncthreads();
    const unsigned int kk = K < nc ? K : nc;
    if (!rq.by_ratio) { if (tid == 0) out_cnt[rq.out] = kk; return; }
    // ratio filter on the sorted list: keep while ln v0 - ln v < ratio (table.rs:134-149); NaN compares false
    __shared__ unsigned int s_keep;
    if (tid == 0) s_keep = 0;
    __syncthreads();
    for (unsigned int r = tid; r < kk; r += SELECT_THREADS) {
        uint32_t g = out_ids[(size_t)rq.out * MAX_ACTIVE + r];
        double L = xlog(xf(gm[g] + gi[g] + gd[g], ge[g]));
        if (s_L0 - L < rq.ratio) atomicAdd(&s_keep, 1u);  // sorted descending => the kept ones are a prefix
    }
    __syncthreads();
    if (tid == 0) out_cnt[rq.out] = s_keep;
}

// ------------------------------------------------------------------------------------------------ host wrappers
static int fast_span(const LinParams& lp) {
    // binades one row can take away from a value: emission/transition factors of round A (3) + the Del chain (6) +
    // an init/trans factor (2^-40 covers N up to 10^12 copies).  The common frame is used only if that still leaves
    // every product a normal f64.
    double pmin = 1.0;
    const double ps[] = {lp.p_mismatch, lp.p_match, lp.p_random, lp.p_end, lp.p_MM, lp.p_IM, lp.p_DM, lp.p_MI, lp.p_II, lp.p_DI, lp.p_MD, lp.p_ID, lp.p_DD};
    for (double p : ps) if (p > 0.0 && p < pmin) pmin = p;
    int l2 = (int)std::ceil(-std::log2(pmin)) + 1;
    int span = 1000 - (9 * l2 + 80);
    if (const char* f = getenv("DBGPHMM_FORCE_EXACT")) if (f[0] == '1') return -1;
    return span < 64 ? -1 : (span > 900 ? 900 : span);
}

// reads handled by one CTA: the tile structure is staged once per CTA, so as many as still leave ~12 waves of CTAs -- and among the
// next few splits the one whose last, partly filled wave costs least: the CTAs of a launch take the same time, so a grid of 12.2
// waves runs as long as one of 13.  Cost model: ceil(waves) x (jobs per CTA + 3), the 3 standing for the per-CTA staging of the tile.
static uint32_t fast_jobs_per_cta(const dbgphmm_model* m, uint32_t n_chunks, uint32_t n_jobs) {
    const uint32_t target = 8u * 3u * (uint32_t)m->n_sm;
    const uint32_t n_ctas = (n_chunks + WT_WARPS - 1) / WT_WARPS;
    const uint32_t resident = (uint32_t)WT_MIN_CTAS * (uint32_t)m->n_sm;
    const uint32_t g0 = std::max<uint32_t>(1, std::min<uint32_t>(n_jobs, (target + n_ctas - 1) / n_ctas));
    uint32_t best_jpc = (n_jobs + g0 - 1) / g0;
    uint64_t best_cost = ~0ull;
    for (uint32_t g = g0; g <= std::min<uint32_t>(n_jobs, g0 + 5); g++) {
        const uint32_t jpc = (n_jobs + g - 1) / g, groups = (n_jobs + jpc - 1) / jpc;
        const uint64_t waves = ((uint64_t)n_ctas * groups + resident - 1) / resident;
        const uint64_t cost = waves * (jpc + 3);
        if (cost < best_cost) { best_cost = cost; best_jpc = jpc; }
    }
    return std::max<uint32_t>(1, best_jpc);
}

int dense_configure(dbgphmm_model* m) {
    CUDA_TRY(cudaFuncSetAttribute(k_dense_fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, DENSE_SMEM_BYTES));
    CUDA_TRY(cudaFuncSetAttribute(k_dense_bwd, cudaFuncAttributeMaxDynamicSharedMemorySize, DENSE_SMEM_BYTES));
    CUDA_TRY(cudaFuncSetAttribute(k_dense_reg<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, RS_SMEM_BYTES));
    CUDA_TRY(cudaFuncSetAttribute(k_dense_reg<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, RS_SMEM_BYTES));
    CUDA_TRY(cudaFuncSetAttribute(k_dense_fwd2, cudaFuncAttributeMaxDynamicSharedMemorySize, RS2_SMEM_BYTES));
    CUDA_TRY(cudaFuncSetAttribute(k_dense_bwd2, cudaFuncAttributeMaxDynamicSharedMemorySize, RS2_SMEM_BYTES));
    CUDA_TRY(cudaFuncSetAttribute(k_dense_select, cudaFuncAttributeMaxDynamicSharedMemorySize, SELECT_SMEM_BYTES));
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, m->device));
    m->n_sm = prop.multiProcessorCount;
    return DBGPHMM_OK;
}

static int ensure_jstep(dbgphmm_model* m, uint32_t n_jobs) {
    if (n_jobs <= MSET(m).jstep_cap) return DBGPHMM_OK;
    CUDA_TRY(cudaStreamSynchronize(MSET(m).stream));
    cudaFree(MSET(m).d_jstep); MSET(m).d_jstep = nullptr; MSET(m).jstep_cap = 0;
    const uint32_t cap = std::max<uint32_t>(1024, n_jobs + n_jobs / 2);
    CUDA_TRY(cudaMalloc(&MSET(m).d_jstep, (size_t)cap * sizeof(JStep)));
    MSET(m).jstep_cap = cap;
    return DBGPHMM_OK;
}

// worklist: [0] = number of tiles left to the exact kernel, [1..] = (job << 32 | chunk)
int dense_forward_step(dbgphmm_model* m, const DensePool& pool, const DJob* d_jobs, uint32_t n_jobs, uint32_t s,
                       const uint8_t* d_bases, RowDesc* d_desc, const int* d_active, XF* d_partials, unsigned long long* d_worklist,
                       uint64_t step_cells) {
    CUDA_TRY(cudaMemsetAsync(d_worklist, 0, sizeof(unsigned long long), MSET(m).stream));
    const uint32_t jpc = fast_jobs_per_cta(m, m->fwd.n_chunks, n_jobs);
    dim3 grid((m->fwd.n_chunks + WT_WARPS - 1) / WT_WARPS, (n_jobs + jpc - 1) / jpc);
    ST_TRY(ensure_jstep(m, n_jobs));
    k_dense_prep<true><<<(n_jobs + 127) / 128, 128, 0, MSET(m).stream>>>(m->lin, d_jobs, n_jobs, s, d_bases, d_desc, d_active, pool.base, pool.slab_bytes, (JStep*)MSET(m).d_jstep);
    COUNT_LAUNCH();
    launch_timer_begin(MSET(m).stream);
    k_dense_reg<true><<<grid, WT_WARPS * 32, RS_SMEM_BYTES, MSET(m).stream>>>(plan_view(m->fwd), graph_view(m), m->lin, (const JStep*)MSET(m).d_jstep, n_jobs, pool.Np,
                                                                          d_partials, m->fwd.n_chunks, fast_span(m->lin), d_worklist, jpc);
    COUNT_LAUNCH();
    k_dense_fwd<<<16 * m->n_sm, DENSE_THREADS, DENSE_SMEM_BYTES, MSET(m).stream>>>(plan_view(m->fwd), graph_view(m), m->lin, d_jobs, s, d_bases, d_desc,
                                                                             d_active, pool.base, pool.slab_bytes, pool.Np, d_partials, m->fwd.n_chunks, d_worklist);
    launch_timer_end(MSET(m).stream, step_cells);
    COUNT_LAUNCH();
    k_dense_fwd_finish<<<n_jobs, 256, 0, MSET(m).stream>>>(m->lin, d_jobs, s, d_desc, d_active, d_partials, m->fwd.n_chunks);
    COUNT_LAUNCH();
    CUDA_TRY(cudaGetLastError());
    return DBGPHMM_OK;
}

// span of the two-rows-per-launch kernel: room for the factors of two rows
static int fast_span2(const LinParams& lp) {
    double pmin = 1.0;
    const double ps[] = {lp.p_mismatch, lp.p_match, lp.p_random, lp.p_end, lp.p_MM, lp.p_IM, lp.p_DM, lp.p_MI, lp.p_II, lp.p_DI, lp.p_MD, lp.p_ID, lp.p_DD};
    for (double p : ps) if (p > 0.0 && p < pmin) pmin = p;
    int l2 = (int)std::ceil(-std::log2(pmin)) + 1;
    int span = 1000 - 2 * (9 * l2 + 80);
    if (const char* f = getenv("DBGPHMM_FORCE_EXACT")) if (f[0] == '1') return -1;
    if (const char* f = getenv("DBGPHMM_DENSE_SPAN2")) { int v = atoi(f); if (v > 0 && v < span) return v; }   // tests: force the fallback
    return span < 64 ? -1 : (span > 900 ? 900 : span);
}
bool dense_can_pair(const dbgphmm_model* m) { return m->fwd2.n_chunks > 0 && m->bwd2.n_chunks > 0 && fast_span2(m->lin) > 0; }
uint32_t dense_pair_tiles(const dbgphmm_model* m, int dir) { return dir == 0 ? m->fwd2.n_chunks : m->bwd2.n_chunks; }

// backward rows first_row - s and first_row - s - 1 of every job in one launch ; d_partials holds [2][n_jobs][bwd2.n_chunks][2]
int dense_backward_pair(dbgphmm_model* m, const DensePool& pool, const DJob* d_jobs, uint32_t n_jobs, uint32_t s, const uint8_t* d_bases,
                        RowDesc* d_desc, XF* d_partials, int* d_redo, uint64_t pair_cells, bool second) {
    const uint32_t nt = m->bwd2.n_chunks;
    const uint32_t jpc = fast_jobs_per_cta(m, nt, n_jobs);
    dim3 grid((nt + WT_WARPS - 1) / WT_WARPS, (n_jobs + jpc - 1) / jpc);
    ST_TRY(ensure_jstep(m, 2 * n_jobs));
    k_dense_prep2<false><<<(n_jobs + 127) / 128, 128, 0, MSET(m).stream>>>(m->lin, d_jobs, n_jobs, s, d_bases, d_desc, pool.base, pool.slab_bytes, (JStep*)MSET(m).d_jstep);
    COUNT_LAUNCH();
    launch_timer_begin(MSET(m).stream);
    k_dense_bwd2<<<grid, WT_WARPS * 32, RS2_SMEM_BYTES, MSET(m).stream>>>(plan_view(m->bwd2), graph_view(m), m->lin, (const JStep*)MSET(m).d_jstep, n_jobs, pool.Np,
                                                                      d_partials, nt, fast_span2(m->lin), d_redo, jpc);
    launch_timer_end(MSET(m).stream, pair_cells);
    COUNT_LAUNCH();
    k_dense_bwd_finish<<<n_jobs, 256, 0, MSET(m).stream>>>(m->lin, d_jobs, s, d_desc, nullptr, d_partials, nt, 1);
    COUNT_LAUNCH();
    if (second) {
        k_dense_bwd_finish<<<n_jobs, 256, 0, MSET(m).stream>>>(m->lin, d_jobs, s + 1, d_desc, nullptr, d_partials + (size_t)n_jobs * nt * 2, nt, 1);
        COUNT_LAUNCH();
    }
    CUDA_TRY(cudaGetLastError());
    return DBGPHMM_OK;
}

// backward step restricted to the (job, tile) pairs of a prebuilt worklist (recompute pass of the stream strategy), no row reduction
int dense_backward_step_list(dbgphmm_model* m, const DensePool& pool, const DJob* d_jobs, uint32_t s, const uint8_t* d_bases, XF* d_partials,
                             const unsigned long long* d_worklist) {
    k_dense_bwd<<<16 * m->n_sm, DENSE_THREADS, DENSE_SMEM_BYTES, MSET(m).stream>>>(plan_view(m->bwd), graph_view(m), m->lin, d_jobs, s, d_bases, nullptr,
                                                                             pool.base, pool.slab_bytes, pool.Np, d_partials, m->bwd.n_chunks, d_worklist);
    COUNT_LAUNCH();
    CUDA_TRY(cudaGetLastError());
    return DBGPHMM_OK;
}

// forward rows s and s + 1 of every job in one launch (see k_dense_fwd2) ; d_partials holds [2][n_jobs][fwd2.n_chunks]
int dense_forward_pair(dbgphmm_model* m, const DensePool& pool, const DJob* d_jobs, uint32_t n_jobs, uint32_t s, const uint8_t* d_bases,
                       RowDesc* d_desc, XF* d_partials, int* d_redo, uint64_t pair_cells, bool second) {
    const uint32_t nt = m->fwd2.n_chunks;
    const uint32_t jpc = fast_jobs_per_cta(m, nt, n_jobs);
    dim3 grid((nt + WT_WARPS - 1) / WT_WARPS, (n_jobs + jpc - 1) / jpc);
    ST_TRY(ensure_jstep(m, 2 * n_jobs));
    k_dense_prep2<true><<<(n_jobs + 127) / 128, 128, 0, MSET(m).stream>>>(m->lin, d_jobs, n_jobs, s, d_bases, d_desc, pool.base, pool.slab_bytes, (JStep*)MSET(m).d_jstep);
    COUNT_LAUNCH();
    launch_timer_begin(MSET(m).stream);
    k_dense_fwd2<<<grid, WT_WARPS * 32, RS2_SMEM_BYTES, MSET(m).stream>>>(plan_view(m->fwd2), graph_view(m), m->lin, (const JStep*)MSET(m).d_jstep, n_jobs, pool.Np,
                                                                      d_partials, nt, fast_span2(m->lin), d_redo, jpc);
    launch_timer_end(MSET(m).stream, pair_cells);
    COUNT_LAUNCH();
    k_dense_fwd_finish<<<n_jobs, 256, 0, MSET(m).stream>>>(m->lin, d_jobs, s, d_desc, nullptr, d_partials, nt, 1);
    COUNT_LAUNCH();
    if (second) {
        k_dense_fwd_finish<<<n_jobs, 256, 0, MSET(m).stream>>>(m->lin, d_jobs, s + 1, d_desc, nullptr, d_partials + (size_t)n_jobs * nt, nt, 1);
        COUNT_LAUNCH();
    }
    CUDA_TRY(cudaGetLastError());
    return DBGPHMM_OK;
}

int dense_forward_step_list(dbgphmm_model* m, const DensePool& pool, const DJob* d_jobs, uint32_t s, const uint8_t* d_bases,
                            const RowDesc* d_desc, XF* d_partials, const unsigned long long* d_worklist) {
    k_dense_fwd<<<16 * m->n_sm, DENSE_THREADS, DENSE_SMEM_BYTES, MSET(m).stream>>>(plan_view(m->fwd), graph_view(m), m->lin, d_jobs, s, d_bases, d_desc,
                                                                             nullptr, pool.base, pool.slab_bytes, pool.Np, d_partials, m->fwd.n_chunks, d_worklist);
    COUNT_LAUNCH();
    CUDA_TRY(cudaGetLastError());
    return DBGPHMM_OK;
}

int dense_backward_step(dbgphmm_model* m, const DensePool& pool, const DJob* d_jobs, uint32_t n_jobs, uint32_t s,
                        const uint8_t* d_bases, RowDesc* d_desc, const int* d_active, XF* d_partials, unsigned long long* d_worklist,
                        uint64_t step_cells) {
    CUDA_TRY(cudaMemsetAsync(d_worklist, 0, sizeof(unsigned long long), MSET(m).stream));
    const uint32_t jpc = fast_jobs_per_cta(m, m->bwd.n_chunks, n_jobs);
    dim3 grid((m->bwd.n_chunks + WT_WARPS - 1) / WT_WARPS, (n_jobs + jpc - 1) / jpc);
    ST_TRY(ensure_jstep(m, n_jobs));
    k_dense_prep<false><<<(n_jobs + 127) / 128, 128, 0, MSET(m).stream>>>(m->lin, d_jobs, n_jobs, s, d_bases, d_desc, d_active, pool.base, pool.slab_bytes, (JStep*)MSET(m).d_jstep);
    COUNT_LAUNCH();
    launch_timer_begin(MSET(m).stream);
    k_dense_reg<false><<<grid, WT_WARPS * 32, RS_SMEM_BYTES, MSET(m).stream>>>(plan_view(m->bwd), graph_view(m), m->lin, (const JStep*)MSET(m).d_jstep, n_jobs, pool.Np,
                                                                           d_partials, m->bwd.n_chunks, fast_span(m->lin), d_worklist, jpc);
    COUNT_LAUNCH();
    k_dense_bwd<<<16 * m->n_sm, DENSE_THREADS, DENSE_SMEM_BYTES, MSET(m).stream>>>(plan_view(m->bwd), graph_view(m), m->lin, d_jobs, s, d_bases, d_active,
                                                                             pool.base, pool.slab_bytes, pool.Np, d_partials, m->bwd.n_chunks, d_worklist);
    launch_timer_end(MSET(m).stream, step_cells);
    COUNT_LAUNCH();
    k_dense_bwd_finish<<<n_jobs, 256, 0, MSET(m).stream>>>(m->lin, d_jobs, s, d_desc, d_active, d_partials, m->bwd.n_chunks);
    COUNT_LAUNCH();
    CUDA_TRY(cudaGetLastError());
    return DBGPHMM_OK;
}

int dense_select(dbgphmm_model* m, const DensePool& pool, const SelectReq* d_reqs, uint32_t n_reqs, const int* d_active,
                 uint32_t* d_out_ids, uint32_t* d_out_cnt) {
    if (n_reqs == 0) return DBGPHMM_OK;
    // rows of >= 64 K cells: tile maxima first (one streaming pass), the selection then reads ~k tiles instead of the row
    const uint32_t n_tiles = (m->N + SEL_TILE - 1) / SEL_TILE;
    bool pre = m->N >= (1u << 16);
    if (const char* e = getenv("DBGPHMM_SELECT_TILES")) pre = e[0] == '1';
    DevBuf b_tiles;
    if (pre && b_tiles.alloc(sizeof(TileKey) * (size_t)n_reqs * n_tiles) != DBGPHMM_OK) pre = false;   // (no scratch left: the full sweep needs none)
    if (pre) {
        k_select_tilemax<<<dim3((n_tiles + 7) / 8, n_reqs), 256, 0, MSET(m).stream>>>(d_reqs, d_active, pool.base, pool.slab_bytes, pool.Np, m->N, n_tiles, b_tiles.as<TileKey>());
        COUNT_LAUNCH();
        CUDA_TRY(cudaGetLastError());
    }
    k_dense_select<<<n_reqs, SELECT_THREADS, SELECT_SMEM_BYTES, MSET(m).stream>>>(graph_view(m), d_reqs, d_active, pool.base, pool.slab_bytes, pool.Np, d_out_ids, d_out_cnt,
                                                                            pre ? b_tiles.as<TileKey>() : nullptr, n_tiles);
    COUNT_LAUNCH();
    CUDA_TRY(cudaGetLastError());
    return DBGPHMM_OK;
}
