// products.cu — quantities that need forward AND backward rows (PHMMOutput, table.rs:450-517):
//   to_emit_probs(i) = F[i] (x) B[i] / P                       table.rs:500-505
//   to_state_probs / to_node_freqs = sum_i, merge m+i+d, exp    freq.rs:237-255
//   to_mapping / to_mapping_by_score_ratio                      hint.rs:124-142
// Merged index t = 1..n pairs forward row t-1 with backward row t (row n = b_init, p_end everywhere,
// table.rs:414-434, backward.rs:197-211).  t = 0 contributes nothing to node states (f_init has m=i=d=0).
#include <algorithm>
#include "engine.h"

struct ProdCtx {
    const RowDesc *fdesc, *bdesc;
    const uint64_t *fdesc0, *bdesc0;
    const uint32_t* len;
    const char *farena, *barena, *fpool, *bpool;
    uint64_t fslab_bytes, bslab_bytes;
    uint32_t fNp, bNp;
    const XF* P;  // per job: forward full prob
    int f_dense_ok, b_dense_ok;  // 0: the dense rows of that direction were not kept (their pairs are taken on the fly)
    const uint32_t* orig_of;
    double p_end;
    uint32_t N;
};

struct RowView {  // one stored row
    int kind;     // ROW_DENSE / ROW_SPARSE / 3 = b_init
    uint32_t n_ent, n_mi;
    const double *m, *i, *d; const uint32_t* id; const int* ex;
};
__device__ __forceinline__ RowView view_row(const RowDesc& r, const char* arena, const char* pool, uint64_t slab_bytes, uint32_t Np) {
    RowView v; v.kind = r.kind; v.n_ent = r.n_ent; v.n_mi = r.n_mi; v.id = nullptr;
    if (r.kind == ROW_DENSE) {
        const char* sl = pool + r.off * slab_bytes;
        v.m = (const double*)sl; v.i = v.m + Np; v.d = v.i + Np; v.ex = (const int*)(v.d + Np);
    } else {
        const char* pay = arena + r.off;
        v.m = (const double*)pay; v.i = v.m + r.n_ent; v.d = v.i + r.n_ent; v.id = (const uint32_t*)(v.d + r.n_ent); v.ex = (const int*)(v.id + r.n_ent);
    }
    return v;
}

#define PROD_MAXE 832
// value of row `v` at node `id` (dense / init / sparse with ids staged in shared memory)
__device__ __forceinline__ void row_at(const RowView& v, const uint32_t* sh_ids, uint32_t id, double p_end, double* m, double* i, double* d, int* ex) {
    if (v.kind == ROW_DENSE) { *m = v.m[id]; *i = v.i[id]; *d = v.d[id]; *ex = v.ex[id]; return; }
    if (v.kind == 3) { *m = *i = *d = p_end; *ex = 0; return; }
    for (uint32_t e = 0; e < v.n_ent; e++)
        if (sh_ids[e] == id) { *m = v.m[e]; *i = v.i[e]; *d = v.d[e]; *ex = v.ex[e]; return; }
    *m = *i = *d = 0.0; *ex = 0;
}

// ---- node frequencies, rows where at least one side is sparse.  One WARP per merged row (a sparse row holds ~50 entries), four
// rows per CTA: a CTA per row spent most of the launch on CTA scheduling (10^7 CTAs of one busy warp each on the C3 workload).
#define PROD_ROWS_PER_CTA 4
__global__ void k_prod_rows_freq(ProdCtx C, double* __restrict__ freq, int* __restrict__ err) {
    const uint32_t j = blockIdx.y, lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const uint32_t r = blockIdx.x * PROD_ROWS_PER_CTA + w;
    if (r >= C.len[j]) return;   // (whole warps leave ; no block-wide barrier below)
    const XF P = C.P[j];
    if (P.v == 0.0) { if (lane == 0) *err = 1; return; }
    const RowDesc fr = C.fdesc[C.fdesc0[j] + r];
    RowDesc br;
    if (r + 1 < C.len[j]) br = C.bdesc[C.bdesc0[j] + r + 1]; else { br.kind = 3; br.n_ent = 0; br.n_mi = 0; br.n_d = 0; br.off = 0; }
    if (fr.kind == ROW_DENSE && br.kind != ROW_SPARSE) return;  // dense x dense: k_prod_dense_freq
    if ((fr.kind == ROW_DENSE && !C.f_dense_ok) || (br.kind == ROW_DENSE && !C.b_dense_ok)) return;  // taken by step_products
    const RowView F = view_row(fr, C.farena, C.fpool, C.fslab_bytes, C.fNp);
    const RowView B = view_row(br, C.barena, C.bpool, C.bslab_bytes, C.bNp);
    __shared__ uint32_t sh_all[PROD_ROWS_PER_CTA][PROD_MAXE];
    uint32_t* sh_ids = sh_all[w];
    const RowView& it = (fr.kind == ROW_SPARSE) ? F : B;      // iterate the sparse operand (lhs if both sparse)
    const RowView& other = (fr.kind == ROW_SPARSE) ? B : F;
    if (other.kind == ROW_SPARSE) for (uint32_t e = lane; e < other.n_ent && e < PROD_MAXE; e += 32) sh_ids[e] = other.id[e];
    __syncwarp();
    for (uint32_t e = lane; e < it.n_ent; e += 32) {
        uint32_t id = it.id[e];
        double m2, i2, d2; int e2;
        row_at(other, sh_ids, id, C.p_end, &m2, &i2, &d2, &e2);
        double v = it.m[e] * m2 + it.i[e] * i2 + it.d[e] * d2;
        if (v > 0.0) {
            double w2 = (v / P.v) * pow2i(it.ex[e] + e2 - P.e);
            if (w2 > 0.0) atomicAdd(&freq[C.orig_of[id]], w2);
        }
    }
}

// ---- edge and init frequencies (PHMMOutput::to_edge_and_init_freqs, freq.rs:276-298 over to_trans_and_init_probs :332-389)
// Merged index t = 0..n (one CTA each): F = forward merged(t) (t == 0: f_init), B2 = backward merged(t + 1) (used when t < n),
// B1 = backward merged(t).  Per edge (k -> l):
//   to Match:  a_kl e_l(x[t]) B2.m[l] (p_MM F.m[k] + p_IM F.i[k] + p_DM F.d[k]) / P        (t < n)
//   to Del:    a_kl           B1.d[l] (p_MD F.m[k] + p_ID F.i[k] + p_DD F.d[k]) / P
// and per node v the Begin -> v transitions with (F.mb, F.ib) and init_v in place of (F.m, F.i) and a_kl.  The two kinds of terms
// are accumulated separately (the reference sums the six terms of one (t, e) in log space first; the difference is rounding).
// Each pair of rows is walked from its sparser side: entries of a sparse backward row and their parents, else entries of the
// forward row (sparse, or all N nodes) and their children.
struct EdgeCtx {
    const uint32_t *par_off, *par_node, *par_eid, *chi_off, *chi_node, *chi_eid;
    const uint8_t* emission;
    const double *init, *trans;
    const uint8_t* bases;      // the read of job j starts at base_off[j]
    const uint64_t* base_off;
    LinParams lp;
};
__device__ __forceinline__ void edge_add(double* acc, double v, int e, const XF& P) {
    if (v > 0.0) { const double w = (v / P.v) * pow2i(e - P.e); if (w > 0.0) atomicAdd(acc, w); }
}
// one (forward row, backward row) pair ; to_m selects the Match terms (backward m, emission) or the Del terms (backward d)
__device__ void edge_pair(const ProdCtx& C, const EdgeCtx& G, const RowView& F, bool f_is_init, XF fmb, XF fib, const RowView& B, bool to_m, uint8_t x,
                          const XF& P, const uint32_t* f_ids, const uint32_t* b_ids, double* edge, double* init) {
    const LinParams& lp = G.lp;
    const double cm = to_m ? lp.p_MM : lp.p_MD, ci = to_m ? lp.p_IM : lp.p_ID, cd = to_m ? lp.p_DM : lp.p_DD;
    auto bval = [&](uint32_t l, double* v, int* e) {   // backward factor of node l
        double m, i, d; int ex;
        row_at(B, b_ids, l, C.p_end, &m, &i, &d, &ex);
        *v = to_m ? m * (G.emission[l] == x ? lp.p_match : lp.p_mismatch) : d; *e = ex;
    };
    // ---- Begin -> v
    {
        const XF bsc = xadd(xmul(fmb, cm), xmul(fib, ci));
        if (bsc.v != 0.0) {
            if (B.kind == ROW_SPARSE) {
                for (uint32_t q = threadIdx.x; q < B.n_ent; q += blockDim.x) {
                    const uint32_t v = B.id[q];
                    const double b = to_m ? B.m[q] * (G.emission[v] == x ? lp.p_match : lp.p_mismatch) : B.d[q];
                    edge_add(&init[C.orig_of[v]], bsc.v * G.init[v] * b, bsc.e + B.ex[q], P);
                }
            } else {
                for (uint32_t v = threadIdx.x; v < C.N; v += blockDim.x) {
                    double b; int be; bval(v, &b, &be);
                    edge_add(&init[C.orig_of[v]], bsc.v * G.init[v] * b, bsc.e + be, P);
                }
            }
        }
    }
    if (f_is_init) return;   // f_init has m = i = d = 0
    // ---- k -> l
    if (B.kind == ROW_SPARSE) {
        for (uint32_t q = threadIdx.x; q < B.n_ent; q += blockDim.x) {
            const uint32_t l = B.id[q];
            const double b = to_m ? B.m[q] * (G.emission[l] == x ? lp.p_match : lp.p_mismatch) : B.d[q];
            if (b == 0.0) continue;
            for (uint32_t a = G.par_off[l]; a < G.par_off[l + 1]; a++) {
                double m, i, d; int fe;
                row_at(F, f_ids, G.par_node[a], 0.0, &m, &i, &d, &fe);
                edge_add(&edge[G.par_eid[a]], G.trans[G.par_eid[a]] * b * (cm * m + ci * i + cd * d), fe + B.ex[q], P);
            }
        }
    } else {
        const uint32_t nf = F.kind == ROW_SPARSE ? F.n_ent : C.N;
        for (uint32_t q = threadIdx.x; q < nf; q += blockDim.x) {
            const uint32_t k = F.kind == ROW_SPARSE ? F.id[q] : q;
            const double f = cm * F.m[q] + ci * F.i[q] + cd * F.d[q];
            if (f == 0.0) continue;
            for (uint32_t a = G.chi_off[k]; a < G.chi_off[k + 1]; a++) {
                double b; int be; bval(G.chi_node[a], &b, &be);
                edge_add(&edge[G.chi_eid[a]], G.trans[G.chi_eid[a]] * b * f, F.ex[q] + be, P);
            }
        }
    }
}
__global__ void k_prod_edge_freq(ProdCtx C, EdgeCtx G, double* __restrict__ edge, double* __restrict__ init, int* __restrict__ err) {
    const uint32_t j = blockIdx.y, t = blockIdx.x, n = C.len[j];
    if (t > n) return;
    const XF P = C.P[j];
    if (P.v == 0.0) { if (threadIdx.x == 0) *err = 1; return; }
    RowDesc fr, b1, b2;
    const bool f_is_init = t == 0;
    if (f_is_init) { fr.kind = ROW_SPARSE; fr.n_ent = 0; fr.n_mi = 0; fr.n_d = 0; fr.off = 0; fr.mb = xf(1.0, 0); fr.ib = xf_zero(); }
    else fr = C.fdesc[C.fdesc0[j] + t - 1];
    auto brow = [&](uint32_t r) { RowDesc d; if (r >= n) { d.kind = 3; d.n_ent = 0; d.n_mi = 0; d.n_d = 0; d.off = 0; } else d = C.bdesc[C.bdesc0[j] + r]; return d; };
    b1 = brow(t); b2 = brow(t + 1);
    if ((fr.kind == ROW_DENSE && !C.f_dense_ok) || ((b1.kind == ROW_DENSE || b2.kind == ROW_DENSE) && !C.b_dense_ok)) { if (threadIdx.x == 0) *err = 2; return; }
    const RowView F = view_row(fr, C.farena, C.fpool, C.fslab_bytes, C.fNp);
    const RowView B1 = view_row(b1, C.barena, C.bpool, C.bslab_bytes, C.bNp);
    const RowView B2 = view_row(b2, C.barena, C.bpool, C.bslab_bytes, C.bNp);
    __shared__ uint32_t f_ids[PROD_MAXE], b_ids[PROD_MAXE];
    if (F.kind == ROW_SPARSE) for (uint32_t e = threadIdx.x; e < F.n_ent && e < PROD_MAXE; e += blockDim.x) f_ids[e] = F.id[e];
    if (t < n) {
        if (B2.kind == ROW_SPARSE) for (uint32_t e = threadIdx.x; e < B2.n_ent && e < PROD_MAXE; e += blockDim.x) b_ids[e] = B2.id[e];
        __syncthreads();
        edge_pair(C, G, F, f_is_init, fr.mb, fr.ib, B2, true, G.bases[G.base_off[j] + t], P, f_ids, b_ids, edge, init);
        __syncthreads();
    }
    if (B1.kind == ROW_SPARSE) for (uint32_t e = threadIdx.x; e < B1.n_ent && e < PROD_MAXE; e += blockDim.x) b_ids[e] = B1.id[e];
    __syncthreads();
    edge_pair(C, G, F, f_is_init, fr.mb, fr.ib, B1, false, 0, P, f_ids, b_ids, edge, init);
}

struct DensePair { uint64_t fslab; uint64_t bslab; int b_init; uint32_t job; };
__global__ void k_prod_dense_freq(ProdCtx C, const DensePair* __restrict__ pairs, double* __restrict__ freq, int* __restrict__ err) {
    const DensePair pr = pairs[blockIdx.y];
    const XF P = C.P[pr.job];
    if (P.v == 0.0) { if (threadIdx.x == 0 && blockIdx.x == 0) *err = 1; return; }
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= C.N) return;
    const char* fs = C.fpool + pr.fslab * C.fslab_bytes;
    const double* fm = (const double*)fs; const double* fi = fm + C.fNp; const double* fd = fi + C.fNp; const int* fe = (const int*)(fd + C.fNp);
    double v; int e2;
    if (pr.b_init) { v = (fm[g] + fi[g] + fd[g]) * C.p_end; e2 = 0; }
    else {
        const char* bs = C.bpool + pr.bslab * C.bslab_bytes;
        const double* bm = (const double*)bs; const double* bi = bm + C.bNp; const double* bd = bi + C.bNp; const int* be = (const int*)(bd + C.bNp);
        v = fm[g] * bm[g] + fi[g] * bi[g] + fd[g] * bd[g]; e2 = be[g];
    }
    if (v > 0.0) {
        double w = (v / P.v) * pow2i(fe[g] + e2 - P.e);
        if (w > 0.0) atomicAdd(&freq[C.orig_of[g]], w);
    }
}


// ---- mapping rows where at least one side is sparse: top nodes of the merged emit probabilities
// pass 0 writes counts[lin]; pass 1 writes (original node id, ln prob) at offsets[lin].
__global__ void k_map_rows(ProdCtx C, int by_ratio, uint32_t n_active, double ratio, int pass, const uint64_t* __restrict__ lin0,
                           uint32_t* __restrict__ counts, const uint64_t* __restrict__ offsets, uint32_t* __restrict__ out_nodes,
                           double* __restrict__ out_logp, int* __restrict__ err) {
    const uint32_t j = blockIdx.y, r = blockIdx.x;
    if (r >= C.len[j]) return;
    const XF P = C.P[j];
    if (P.v == 0.0) { if (threadIdx.x == 0) *err = 1; return; }
    const uint64_t lin = lin0[j] + r;
    const RowDesc fr = C.fdesc[C.fdesc0[j] + r];
    RowDesc br;
    if (r + 1 < C.len[j]) br = C.bdesc[C.bdesc0[j] + r + 1]; else { br.kind = 3; br.n_ent = 0; br.n_mi = 0; br.n_d = 0; br.off = 0; }
    if (fr.kind == ROW_DENSE && br.kind != ROW_SPARSE) return;  // dense x dense handled through dense_select
    if ((fr.kind == ROW_DENSE && !C.f_dense_ok) || (br.kind == ROW_DENSE && !C.b_dense_ok)) return;
    const RowView F = view_row(fr, C.farena, C.fpool, C.fslab_bytes, C.fNp);
    const RowView B = view_row(br, C.barena, C.bpool, C.bslab_bytes, C.bNp);
    __shared__ uint32_t sh_ids[PROD_MAXE];
    __shared__ int k_T[PROD_MAXE];
    __shared__ unsigned long long k_mant[PROD_MAXE];
    __shared__ double s_L0;
    __shared__ uint32_t s_keep;
    const RowView& it = (fr.kind == ROW_SPARSE) ? F : B;
    const RowView& other = (fr.kind == ROW_SPARSE) ? B : F;
    const uint32_t n = it.n_ent < PROD_MAXE ? it.n_ent : PROD_MAXE;
    if (other.kind == ROW_SPARSE) for (uint32_t e = threadIdx.x; e < other.n_ent && e < PROD_MAXE; e += blockDim.x) sh_ids[e] = other.id[e];
    if (threadIdx.x == 0) { s_L0 = -INFINITY; s_keep = 0; }
    __syncthreads();
    for (uint32_t e = threadIdx.x; e < n; e += blockDim.x) {
        double m2, i2, d2; int e2;
        row_at(other, sh_ids, it.id[e], C.p_end, &m2, &i2, &d2, &e2);
        double v = it.m[e] * m2 + it.i[e] * i2 + it.d[e] * d2;
        if (v == 0.0) { k_T[e] = XF_ZERO_E; k_mant[e] = 0; }
        else { long long b = __double_as_longlong(v); k_T[e] = it.ex[e] + e2 + (int)((b >> 52) & 0x7ff) - 1023; k_mant[e] = (unsigned long long)b & 0xfffffffffffffull; }
    }
    __syncthreads();
    const uint32_t K = by_ratio ? MAX_ACTIVE : n_active;
    const uint32_t KK = K < n ? K : n;
    // ln of the merged emit probability of entry e
    auto logp_of = [&](uint32_t e) -> double {
        if (k_T[e] == XF_ZERO_E) return -INFINITY;
        double mant = __longlong_as_double((long long)(k_mant[e] | 0x3ff0000000000000ull));
        return xlog(xf(mant / P.v, k_T[e] - P.e));
    };
    uint32_t my_rank[(PROD_MAXE + 127) / 128];
    int q = 0;
    for (uint32_t e = threadIdx.x; e < n; e += blockDim.x, q++) {
        int T = k_T[e]; unsigned long long mt = k_mant[e];
        uint32_t rank = 0;
        for (uint32_t f = 0; f < n; f++) {
            int Tf = k_T[f]; unsigned long long mf = k_mant[f];
            rank += ((Tf > T) || (Tf == T && (mf > mt || (mf == mt && f < e)))) ? 1u : 0u;
        }
        my_rank[q] = rank;
        if (rank == 0) s_L0 = logp_of(e);
    }
    __syncthreads();
    q = 0;
    for (uint32_t e = threadIdx.x; e < n; e += blockDim.x, q++) {
        uint32_t rank = my_rank[q];
        if (rank >= KK) continue;
        double L = logp_of(e);
        bool keep = by_ratio ? (s_L0 - L < ratio) : true;
        if (keep) {
            atomicAdd(&s_keep, 1u);
            if (pass == 1) { uint64_t o = offsets[lin] + rank; out_nodes[o] = C.orig_of[it.id[e]]; out_logp[o] = L; }
        }
    }
    __syncthreads();
    if (pass == 0 && threadIdx.x == 0) counts[lin] = s_keep;
}

// dense x dense: write the un-normalised product cell into a scratch slab (selection is scale invariant)
__global__ void k_prod_slab(ProdCtx C, const DensePair* __restrict__ pairs, char* __restrict__ tpool, uint64_t tslab_bytes, uint32_t tNp) {
    const DensePair pr = pairs[blockIdx.y];
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= C.N) return;
    const char* fs = C.fpool + pr.fslab * C.fslab_bytes;
    const double* fm = (const double*)fs; const double* fi = fm + C.fNp; const double* fd = fi + C.fNp; const int* fe = (const int*)(fd + C.fNp);
    char* ts = tpool + (uint64_t)blockIdx.y * tslab_bytes;
    double* tm = (double*)ts; double* ti = tm + tNp; double* td = ti + tNp; int* te = (int*)(td + tNp);
    double a, b, c; int e2 = 0;
    if (pr.b_init) { a = fm[g] * C.p_end; b = fi[g] * C.p_end; c = fd[g] * C.p_end; }
    else {
        const char* bs = C.bpool + pr.bslab * C.bslab_bytes;
        const double* bm = (const double*)bs; const double* bi = bm + C.bNp; const double* bd = bi + C.bNp; const int* be = (const int*)(bd + C.bNp);
        a = fm[g] * bm[g]; b = fi[g] * bi[g]; c = fd[g] * bd[g]; e2 = be[g];
    }
    tm[g] = a; ti[g] = b; td[g] = c; te[g] = fe[g] + e2;
}
__global__ void k_map_dense_emit(ProdCtx C, const DensePair* __restrict__ pairs, const char* __restrict__ tpool, uint64_t tslab_bytes, uint32_t tNp,
                                 const uint32_t* __restrict__ top_ids, const uint32_t* __restrict__ top_cnt, uint32_t* __restrict__ out_nodes,
                                 double* __restrict__ out_logp, int* __restrict__ err) {
    const uint32_t p = blockIdx.x;
    const XF P = C.P[pairs[p].job];
    if (P.v == 0.0) { if (threadIdx.x == 0) *err = 1; return; }
    const char* ts = tpool + (uint64_t)p * tslab_bytes;
    const double* tm = (const double*)ts; const double* ti = tm + tNp; const double* td = ti + tNp; const int* te = (const int*)(td + tNp);
    for (uint32_t r = threadIdx.x; r < top_cnt[p]; r += blockDim.x) {
        uint32_t g = top_ids[(size_t)p * MAX_ACTIVE + r];
        double v = tm[g] + ti[g] + td[g];
        out_nodes[(size_t)p * MAX_ACTIVE + r] = C.orig_of[g];
        out_logp[(size_t)p * MAX_ACTIVE + r] = v == 0.0 ? -INFINITY : xlog(xf(v / P.v, te[g] - P.e));
    }
}

// ---- on-the-fly products of a dense row that lives in a ping-pong slab with the other direction's stored sparse row
__global__ void k_prod_step(const DJob* __restrict__ jobs, uint32_t s, int dir, const char* __restrict__ pool, uint64_t slab_bytes, uint32_t Np,
                            const RowDesc* __restrict__ odesc, const uint64_t* __restrict__ odesc0, const char* __restrict__ oarena,
                            const XF* __restrict__ Pj, const uint32_t* __restrict__ orig_of, double* __restrict__ freq, int* __restrict__ err) {
    const DJob jb = jobs[blockIdx.x];
    if (s >= jb.n_steps) return;
    const int row = dir == 0 ? jb.first_row + (int)s : jb.first_row - (int)s;
    const int orow = dir == 0 ? row + 1 : row - 1;       // forward row r <-> backward row r+1
    if (orow < 0 || orow >= (int)jb.len) return;
    const RowDesc orr = odesc[odesc0[blockIdx.x] + orow];
    if (orr.kind != ROW_SPARSE) return;
    const XF P = Pj[blockIdx.x];
    if (P.v == 0.0) { if (threadIdx.x == 0) *err = 1; return; }
    const uint64_t slab = jb.slab0 + (jb.slab_mod ? (s % jb.slab_mod) : s);
    const char* sl = pool + slab * slab_bytes;
    const double* gm = (const double*)sl; const double* gi = gm + Np; const double* gd = gi + Np; const int* ge = (const int*)(gd + Np);
    const char* pay = oarena + orr.off;
    const double* sm = (const double*)pay; const double* si = sm + orr.n_ent; const double* sd = si + orr.n_ent;
    const uint32_t* sid = (const uint32_t*)(sd + orr.n_ent); const int* sex = (const int*)(sid + orr.n_ent);
    for (uint32_t e = threadIdx.x; e < orr.n_ent; e += blockDim.x) {
        uint32_t id = sid[e];
        double v = sm[e] * gm[id] + si[e] * gi[id] + sd[e] * gd[id];
        if (v > 0.0) {
            double w = (v / P.v) * pow2i(sex[e] + ge[id] - P.e);
            if (w > 0.0) atomicAdd(&freq[orig_of[id]], w);
        }
    }
}
int step_products(dbgphmm_model* m, const StepProducts& sp, const DensePool& pool, const DJob* d_jobs, uint32_t n_jobs, uint32_t s, int dir, uint32_t job0) {
    if (!sp.other->d_desc0) { dbg_set_error("step_products: the other direction has no device row index"); return DBGPHMM_ERR_INVALID; }
    k_prod_step<<<n_jobs, 128, 0, MSET(m).stream>>>(d_jobs, s, dir, pool.base, pool.slab_bytes, pool.Np, sp.other->d_desc, sp.other->d_desc0 + job0,
                                               sp.other->arena.base, sp.P + job0, m->d_orig_of, sp.d_freqs, sp.d_err);
    COUNT_LAUNCH();
    return DBGPHMM_OK;
}

// ------------------------------------------------------------------------------------------------ host side
struct ProdBufs { DevBuf fdesc0, bdesc0, len, pairs, err; };

static int make_ctx(dbgphmm_model* m, const RowStore& F, const RowStore& B, ProdBufs& pb, ProdCtx* C) {
    cudaStream_t st = MSET(m).stream;
    ST_TRY(dev_upload(pb.fdesc0, F.desc0, st)); ST_TRY(dev_upload(pb.bdesc0, B.desc0, st)); ST_TRY(dev_upload(pb.len, F.len, st));
    ST_TRY(pb.err.alloc(sizeof(int)));
    CUDA_TRY(cudaMemsetAsync(pb.err.p, 0, sizeof(int), st));
    C->fdesc = F.d_desc; C->bdesc = B.d_desc; C->fdesc0 = pb.fdesc0.as<uint64_t>(); C->bdesc0 = pb.bdesc0.as<uint64_t>(); C->len = pb.len.as<uint32_t>();
    C->farena = F.arena.base; C->barena = B.arena.base; C->fpool = F.pool.base; C->bpool = B.pool.base;
    C->fslab_bytes = F.pool.slab_bytes; C->bslab_bytes = B.pool.slab_bytes; C->fNp = F.pool.Np; C->bNp = B.pool.Np;
    C->P = F.d_final; C->f_dense_ok = F.dense_kept; C->b_dense_ok = B.dense_kept; C->orig_of = m->d_orig_of; C->p_end = m->lin.p_end; C->N = m->N;
    return DBGPHMM_OK;
}

// (job, forward row r) pairs whose forward row r and backward row r+1 are both dense (or b_init)
static void dense_pairs(const RowStore& F, const RowStore& B, std::vector<DensePair>& pairs, std::vector<uint64_t>* lin, const std::vector<uint64_t>& lin0) {
    if (!F.dense_kept || !B.dense_kept) return;
    for (size_t j = 0; j < F.len.size(); j++) {
        uint32_t n = F.len[j], ndf = F.nd[j];
        int lo = B.bdense_lo[j], hi = B.bdense_hi[j];
        for (uint32_t r = 0; r < ndf && r < n; r++) {
            DensePair p; p.job = (uint32_t)j; p.fslab = F.slab0[j] + r; p.bslab = 0; p.b_init = 0;
            int t = (int)r + 1;
            if (t == (int)n) p.b_init = 1;
            else if (hi >= 0 && t >= lo && t <= hi) p.bslab = B.slab0[j] + (uint64_t)(hi - t);
            else continue;
            pairs.push_back(p);
            if (lin) lin->push_back(lin0[j] + r);
        }
    }
}

int run_products_freqs(dbgphmm_model* m, const std::vector<HJob>& jobs, const RowStore& F, const RowStore& B, double* d_freqs) {
    HostTrace tr("run_products_freqs");
    cudaStream_t st = MSET(m).stream;
    EvTimer tm(st, &g_times.product_ms);
    const uint32_t J = (uint32_t)jobs.size();
    if (J == 0) return DBGPHMM_OK;
    ProdBufs pb; ProdCtx C;
    ST_TRY(make_ctx(m, F, B, pb, &C));
    uint32_t maxlen = 0;
    for (uint32_t j = 0; j < J; j++) maxlen = std::max(maxlen, F.len[j]);
    if (maxlen) {
        dim3 g((maxlen + PROD_ROWS_PER_CTA - 1) / PROD_ROWS_PER_CTA, J);
        k_prod_rows_freq<<<g, 32 * PROD_ROWS_PER_CTA, 0, st>>>(C, d_freqs, pb.err.as<int>()); COUNT_LAUNCH();
    }
    std::vector<DensePair> pairs;
    std::vector<uint64_t> dummy;
    dense_pairs(F, B, pairs, nullptr, dummy);
    if (!pairs.empty()) {
        ST_TRY(dev_upload(pb.pairs, pairs, st));
        for (size_t off = 0; off < pairs.size(); off += 32768) {
            uint32_t cnt = (uint32_t)std::min<size_t>(32768, pairs.size() - off);
            dim3 g((m->N + 255) / 256, cnt);
            k_prod_dense_freq<<<g, 256, 0, st>>>(C, pb.pairs.as<DensePair>() + off, d_freqs, pb.err.as<int>()); COUNT_LAUNCH();
        }
    }
    int err = 0;
    CUDA_TRY(cudaMemcpyAsync(&err, pb.err.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    CUDA_TRY(cudaGetLastError());
    if (err) { dbg_set_error("P(read) == 0: emit probabilities are NaN in the reference (table.rs:500-505)"); return DBGPHMM_ERR_ZERO_PROB; }
    return DBGPHMM_OK;
}

// d_edge[E] (original EdgeIndex order) and d_init[N] (original node order) are accumulated into.  Both directions must have kept
// all their rows (the store strategy).
int run_products_edge_freqs(dbgphmm_model* m, const std::vector<HJob>& jobs, const RowStore& F, const RowStore& B, const uint8_t* d_bases,
                            double* d_edge, double* d_init) {
    cudaStream_t st = MSET(m).stream;
    EvTimer tm(st, &g_times.product_ms);
    const uint32_t J = (uint32_t)jobs.size();
    if (J == 0) return DBGPHMM_OK;
    ProdBufs pb; ProdCtx C;
    ST_TRY(make_ctx(m, F, B, pb, &C));
    std::vector<uint64_t> boff(J);
    uint32_t maxlen = 0;
    for (uint32_t j = 0; j < J; j++) { boff[j] = jobs[j].base_off; maxlen = std::max(maxlen, F.len[j]); }
    DevBuf b_boff;
    ST_TRY(dev_upload(b_boff, boff, st));
    EdgeCtx G{m->d_par_off, m->d_par_node, m->d_par_eid, m->d_chi_off, m->d_chi_node, m->d_chi_eid, m->d_emission,
              m->d_init + (size_t)jobs[0].x * m->N, m->d_trans + (size_t)jobs[0].x * m->E, d_bases, b_boff.as<uint64_t>(), m->lin};
    dim3 g(maxlen + 1, J);
    k_prod_edge_freq<<<g, 128, 0, st>>>(C, G, d_edge, d_init, pb.err.as<int>()); COUNT_LAUNCH();
    int err = 0;
    CUDA_TRY(cudaMemcpyAsync(&err, pb.err.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    CUDA_TRY(cudaGetLastError());
    if (err == 1) { dbg_set_error("P(read) == 0: transition probabilities are NaN in the reference (freq.rs:346)"); return DBGPHMM_ERR_ZERO_PROB; }
    if (err) { dbg_set_error("edge frequencies need the stored rows of both directions"); return DBGPHMM_ERR_INVALID; }
    return DBGPHMM_OK;
}

int run_products_mapping(dbgphmm_model* m, const std::vector<HJob>& jobs, const RowStore& F, const RowStore& B, int by_ratio,
                         uint32_t n_active, double ratio, dbgphmm_mappings* out) {
    HostTrace tr_all("run_products_mapping");
    cudaStream_t st = MSET(m).stream;
    EvTimer tm(st, &g_times.product_ms);
    const uint32_t J = (uint32_t)jobs.size();
    ProdBufs pb; ProdCtx C;
    ST_TRY(make_ctx(m, F, B, pb, &C));
    std::vector<uint64_t> lin0(J + 1, 0);
    uint32_t maxlen = 0;
    for (uint32_t j = 0; j < J; j++) { lin0[j + 1] = lin0[j] + F.len[j]; maxlen = std::max(maxlen, F.len[j]); }
    const uint64_t n_lin = lin0[J];
    DevBuf b_lin0, b_counts, b_offsets, b_nodes, b_logp;
    ST_TRY(dev_upload(b_lin0, lin0, st));
    ST_TRY(b_counts.alloc(sizeof(uint32_t) * std::max<uint64_t>(n_lin, 1)));
    CUDA_TRY(cudaMemsetAsync(b_counts.p, 0, sizeof(uint32_t) * std::max<uint64_t>(n_lin, 1), st));
    std::vector<uint32_t> counts(n_lin, 0);
    std::vector<uint64_t> offsets(n_lin + 1, 0);
    std::vector<uint32_t> s_nodes; std::vector<double> s_logp;
    if (maxlen) {
        dim3 g(maxlen, J);
        k_map_rows<<<g, 128, 0, st>>>(C, by_ratio, n_active, ratio, 0, b_lin0.as<uint64_t>(), b_counts.as<uint32_t>(), nullptr, nullptr, nullptr, pb.err.as<int>());
        COUNT_LAUNCH();
        CUDA_TRY(cudaMemcpyAsync(counts.data(), b_counts.p, sizeof(uint32_t) * n_lin, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        for (uint64_t i = 0; i < n_lin; i++) offsets[i + 1] = offsets[i] + counts[i];
        ST_TRY(dev_upload(b_offsets, offsets, st));
        ST_TRY(b_nodes.alloc(sizeof(uint32_t) * std::max<uint64_t>(offsets[n_lin], 1)));
        ST_TRY(b_logp.alloc(sizeof(double) * std::max<uint64_t>(offsets[n_lin], 1)));
        k_map_rows<<<g, 128, 0, st>>>(C, by_ratio, n_active, ratio, 1, b_lin0.as<uint64_t>(), b_counts.as<uint32_t>(), b_offsets.as<uint64_t>(),
                                      b_nodes.as<uint32_t>(), b_logp.as<double>(), pb.err.as<int>());
        COUNT_LAUNCH();
        s_nodes.resize(offsets[n_lin]); s_logp.resize(offsets[n_lin]);
        if (offsets[n_lin]) {
            CUDA_TRY(cudaMemcpyAsync(s_nodes.data(), b_nodes.p, sizeof(uint32_t) * offsets[n_lin], cudaMemcpyDeviceToHost, st));
            CUDA_TRY(cudaMemcpyAsync(s_logp.data(), b_logp.p, sizeof(double) * offsets[n_lin], cudaMemcpyDeviceToHost, st));
        }
        CUDA_TRY(cudaStreamSynchronize(st));
    }
    // dense x dense rows through the dense top-k selection, in groups bounded by scratch memory
    std::vector<DensePair> pairs; std::vector<uint64_t> plin;
    dense_pairs(F, B, pairs, &plin, lin0);
    std::vector<std::vector<uint32_t>> d_rows_nodes(pairs.size());
    std::vector<std::vector<double>> d_rows_logp(pairs.size());
    if (!pairs.empty()) {
        DensePool tp;
        uint64_t sb = dense_slab_bytes(m->N);
        uint64_t group = std::max<uint64_t>(1, std::min<uint64_t>(pairs.size(), ((uint64_t)2 << 30) / sb));
        tp.Np = (m->N + 1) & ~1u; tp.slab_bytes = sb; tp.n_slabs = group;
        DevBuf b_tp, b_pairs, b_reqs, b_tid, b_tcnt, b_on, b_ol;
        ST_TRY(b_tp.alloc(sb * group)); tp.base = b_tp.as<char>();
        ST_TRY(dev_upload(b_pairs, pairs, st));
        ST_TRY(b_tid.alloc(sizeof(uint32_t) * group * MAX_ACTIVE)); ST_TRY(b_tcnt.alloc(sizeof(uint32_t) * group));
        ST_TRY(b_on.alloc(sizeof(uint32_t) * group * MAX_ACTIVE)); ST_TRY(b_ol.alloc(sizeof(double) * group * MAX_ACTIVE));
        std::vector<SelectReq> reqs(group);
        std::vector<uint32_t> h_cnt(group), h_on(group * MAX_ACTIVE);
        std::vector<double> h_ol(group * MAX_ACTIVE);
        for (uint64_t off = 0; off < pairs.size(); off += group) {
            uint32_t cnt = (uint32_t)std::min<uint64_t>(group, pairs.size() - off);
            dim3 g((m->N + 255) / 256, cnt);
            k_prod_slab<<<g, 256, 0, st>>>(C, b_pairs.as<DensePair>() + off, tp.base, tp.slab_bytes, tp.Np); COUNT_LAUNCH();
            for (uint32_t i = 0; i < cnt; i++) { reqs[i].slab = i; reqs[i].k = by_ratio ? MAX_ACTIVE : n_active; reqs[i].by_ratio = by_ratio; reqs[i].ratio = ratio; reqs[i].active_idx = -1; reqs[i].out = i; }
            ST_TRY(dev_upload(b_reqs, reqs, st));
            ST_TRY(dense_select(m, tp, b_reqs.as<SelectReq>(), cnt, nullptr, b_tid.as<uint32_t>(), b_tcnt.as<uint32_t>()));
            k_map_dense_emit<<<cnt, 128, 0, st>>>(C, b_pairs.as<DensePair>() + off, tp.base, tp.slab_bytes, tp.Np, b_tid.as<uint32_t>(), b_tcnt.as<uint32_t>(),
                                                  b_on.as<uint32_t>(), b_ol.as<double>(), pb.err.as<int>());
            COUNT_LAUNCH();
            CUDA_TRY(cudaMemcpyAsync(h_cnt.data(), b_tcnt.p, sizeof(uint32_t) * cnt, cudaMemcpyDeviceToHost, st));
            CUDA_TRY(cudaMemcpyAsync(h_on.data(), b_on.p, sizeof(uint32_t) * cnt * MAX_ACTIVE, cudaMemcpyDeviceToHost, st));
            CUDA_TRY(cudaMemcpyAsync(h_ol.data(), b_ol.p, sizeof(double) * cnt * MAX_ACTIVE, cudaMemcpyDeviceToHost, st));
            CUDA_TRY(cudaStreamSynchronize(st));
            for (uint32_t i = 0; i < cnt; i++) {
                d_rows_nodes[off + i].assign(h_on.begin() + (size_t)i * MAX_ACTIVE, h_on.begin() + (size_t)i * MAX_ACTIVE + h_cnt[i]);
                d_rows_logp[off + i].assign(h_ol.begin() + (size_t)i * MAX_ACTIVE, h_ol.begin() + (size_t)i * MAX_ACTIVE + h_cnt[i]);
            }
        }
    }
    int err = 0;
    CUDA_TRY(cudaMemcpyAsync(&err, pb.err.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    CUDA_TRY(cudaGetLastError());
    if (err) { dbg_set_error("P(read) == 0: emit probabilities are NaN in the reference (table.rs:500-505)"); return DBGPHMM_ERR_ZERO_PROB; }
    // ---- assemble the host CSR (appending to `out`)
    std::vector<int64_t> dense_of(n_lin, -1);
    for (size_t p = 0; p < plin.size(); p++) dense_of[plin[p]] = (int64_t)p;
    if (out->read_off.empty()) out->read_off.push_back(0);
    if (out->row_off.empty()) out->row_off.push_back(0);
    // (the rows of the sparse pass lie in row order in s_nodes / s_logp: runs of them are appended in one piece -- a C4 batch has
    // 2 M rows of 40 entries, and a vector insert per row was most of the call)
    {
        HostTrace tr_asm("  mapping assembly");
        size_t n_dense_entries = 0;
        for (auto& ns : d_rows_nodes) n_dense_entries += ns.size();
        out->nodes.reserve(out->nodes.size() + s_nodes.size() + n_dense_entries);
        out->logp.reserve(out->logp.size() + s_logp.size() + n_dense_entries);
        out->row_off.reserve(out->row_off.size() + n_lin);
        out->read_off.reserve(out->read_off.size() + J);
        uint64_t run_a = 0, run_b = 0;   // pending run of sparse-pass entries [run_a, run_b)
        auto flush = [&]() {
            if (run_b > run_a) {
                out->nodes.insert(out->nodes.end(), s_nodes.begin() + run_a, s_nodes.begin() + run_b);
                out->logp.insert(out->logp.end(), s_logp.begin() + run_a, s_logp.begin() + run_b);
            }
            run_a = run_b;
        };
        for (uint32_t j = 0; j < J; j++) {
            for (uint32_t r = 0; r < F.len[j]; r++) {
                const uint64_t lin = lin0[j] + r;
                if (dense_of[lin] >= 0) {
                    flush();
                    auto& ns = d_rows_nodes[dense_of[lin]]; auto& ls = d_rows_logp[dense_of[lin]];
                    out->nodes.insert(out->nodes.end(), ns.begin(), ns.end()); out->logp.insert(out->logp.end(), ls.begin(), ls.end());
                    run_a = run_b = offsets[lin + 1];
                    out->row_off.push_back(out->nodes.size());
                } else {
                    if (offsets[lin] != run_b) { flush(); run_a = offsets[lin]; }
                    run_b = offsets[lin + 1];
                    out->row_off.push_back(out->nodes.size() + (run_b - run_a));
                }
            }
            out->read_off.push_back(out->row_off.size() - 1);
        }
        flush();
    }
    return DBGPHMM_OK;
}
