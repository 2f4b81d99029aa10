// mapx.cu — P(read | X) restricted to a Mapping for several candidate parameter sets X at once
// (forward_with_mapping_score_only, forward.rs:79-89, under to_full_prob_reads freq.rs:175-192 / posterior.rs:504-515).
//
// The candidates X of one sampling step share the graph, the read and the mapping: they differ only in init / trans values.
// One CTA therefore walks the rows of one read once for MX_XB candidates: the index work of a row (hash of the row's nodes,
// slots of every parent in the previous and in the current row) is done once, the numeric work runs over (node, candidate)
// pairs with the candidate index fastest, on value arrays laid out [slot][candidate].  Arithmetic per (node, candidate) is
// the same XF sequence as k_sparse in SP_MAPPING mode (non-adaptive: fm, fi, fd0 + 4 x fdt over the row's node list, fe).
// Rows that do not fit (more than MX_CAP nodes, a node with more than MX_PAR parents, a duplicated node) make the group
// fall back to k_sparse.
#include <algorithm>
#include "engine.h"
#include "sparse.h"

#define MX_THREADS 128
// Two shapes of the kernel: the common one -- rows of at most 64 nodes (to_mapping(n_active = 40) rows hold 40, score-ratio rows 4-10),
// 8 candidates per CTA -- and a wide one for everything the reference's 400-entry SparseVec can hold (rows of up to 400 nodes, up to
// 8 parents per node), 2 candidates per CTA, which takes the groups the common shape reports back.
template <int XB_, int CAP_, int PAR_, int HASH_, typename IDX_>
struct MxShape {
    static constexpr int XB = XB_, CAP = CAP_, PAR = PAR_, HASH = HASH_;
    typedef IDX_ idx_t;
    static constexpr unsigned NONE = (unsigned)(idx_t)~(idx_t)0;
};
typedef MxShape<8, 64, 4, 128, uint8_t> MxCommon;
typedef MxShape<2, 400, 8, 1024, uint16_t> MxWide;

struct MJob {
    uint64_t base_off;   // first base of the read
    uint32_t len;
    uint64_t map_row0;   // first row of the read in the mapping CSR
    uint32_t x0, nx;     // candidates x0 .. x0 + nx
    uint32_t out0;       // results at out[out0 + k * out_stride], k < nx
};

template <class SH>
struct MXS {
    double pm[SH::CAP][SH::XB], pi[SH::CAP][SH::XB], pd[SH::CAP][SH::XB];   // previous row, packed cells
    double cm[SH::CAP][SH::XB], ci[SH::CAP][SH::XB], cdv[SH::CAP][SH::XB];  // current row
    double dv[2][SH::CAP][SH::XB];                                           // Del values of the last two rounds
    int pe[SH::CAP][SH::XB], cmie[SH::CAP][SH::XB], cde[SH::CAP][SH::XB], de[2][SH::CAP][SH::XB];
    uint32_t c_id[SH::CAP], par_eid[SH::CAP][SH::PAR];
    uint32_t hkey[2][SH::HASH];
    typename SH::idx_t hval[2][SH::HASH];
    typename SH::idx_t par_prev[SH::CAP][SH::PAR], par_cur[SH::CAP][SH::PAR], self_prev[SH::CAP];
    uint8_t npar[SH::CAP], em_match[SH::CAP];
    XF mb[SH::XB], ib[SH::XB], esum[MX_THREADS / 32][SH::XB];
    int fail;
};

template <int HASH>
__device__ __forceinline__ uint32_t mx_hash(uint32_t id) { return ((id * 2654435761u) >> 12) & (uint32_t)(HASH - 1); }
template <class SH>
__device__ __forceinline__ unsigned mx_find(const uint32_t* key, const typename SH::idx_t* val, uint32_t id) {
    uint32_t h = mx_hash<SH::HASH>(id);
    for (;;) {
        const uint32_t k = key[h];
        if (k == id + 1) return val[h];
        if (k == 0) return SH::NONE;
        h = (h + 1) & (SH::HASH - 1);
    }
}

// (min CTAs per SM: without a residency hint ptxas settled on 32 registers and spilled 72 bytes ; 72 registers, no spills now)
template <class SH>
__global__ void __launch_bounds__(MX_THREADS, SH::CAP <= 64 ? 4 : 2)
k_mapx(const uint32_t* __restrict__ par_off, const uint32_t* __restrict__ par_node, const uint32_t* __restrict__ par_eid,
       const uint8_t* __restrict__ emission, const double* __restrict__ init_t, const double* __restrict__ trans_t, uint32_t n_x,
       LinParams lp, const MJob* __restrict__ jobs, const uint8_t* __restrict__ bases, const uint64_t* __restrict__ map_row_off,
       const uint32_t* __restrict__ map_nodes, XF* __restrict__ out, uint32_t out_stride, int* __restrict__ status,
       unsigned long long* __restrict__ cells) {
    extern __shared__ __align__(16) unsigned char mx_raw[];
    MXS<SH>& S = *(MXS<SH>*)mx_raw;
    constexpr int MX_XB = SH::XB, MX_CAP = SH::CAP, MX_PAR = SH::PAR, MX_HASH = SH::HASH;
    constexpr unsigned MX_NONE = SH::NONE;
    typedef typename SH::idx_t idx_t;
    const MJob jb = jobs[blockIdx.x];
    const int tid = threadIdx.x;
    const int x = tid & (MX_XB - 1), a_lo = tid / MX_XB;          // candidate of this thread ; its nodes are a_lo + 16 j
    const bool x_on = (uint32_t)x < jb.nx;
    const size_t xg = jb.x0 + (x_on ? x : 0);                   // (inactive lanes recompute candidate x0: harmless)
    if (tid < MX_XB) { S.mb[tid] = xf(1.0, 0); S.ib[tid] = xf_zero(); }
    if (tid == 0) S.fail = 0;
    for (int h = tid; h < 2 * MX_HASH; h += MX_THREADS) (&S.hkey[0][0])[h] = 0;
    __syncthreads();
    uint32_t n_prev = 0;
    int cur = 0;             // hash table of the current row
    unsigned long long n_cells = 0;
    XF last = xf_zero();
    for (uint32_t row = 0; row < jb.len; row++) {
        const uint8_t xb = bases[jb.base_off + row];
        const uint64_t r0 = map_row_off[jb.map_row0 + row], r1 = map_row_off[jb.map_row0 + row + 1];
        const uint32_t n = (uint32_t)(r1 - r0);
        if (n > MX_CAP) { if (tid == 0) S.fail = 1; }
        uint32_t* ck = S.hkey[cur]; idx_t* cv = S.hval[cur];
        const uint32_t* pk = S.hkey[cur ^ 1]; const idx_t* pv = S.hval[cur ^ 1];
        // ---- 1. the row's nodes and their hash
        for (int h = tid; h < MX_HASH; h += MX_THREADS) ck[h] = 0;
        __syncthreads();
        if (S.fail) break;
        for (uint32_t a = tid; a < n; a += MX_THREADS) {
            const uint32_t id = map_nodes[r0 + a];
            S.c_id[a] = id;
            uint32_t h = mx_hash<MX_HASH>(id);
            for (;;) {
                const uint32_t old = atomicCAS(&ck[h], 0u, id + 1);
                if (old == 0u) { cv[h] = (idx_t)a; break; }
                if (old == id + 1) { S.fail = 1; break; }   // duplicated node: first-seen semantics live in k_sparse
                h = (h + 1) & (MX_HASH - 1);
            }
        }
        __syncthreads();
        // ---- 2. parents: slots in the previous and in the current row
        for (uint32_t a = tid; a < n; a += MX_THREADS) {
            const uint32_t id = S.c_id[a];
            const uint32_t po = par_off[id], np = par_off[id + 1] - po;
            if (np > MX_PAR) S.fail = 1;
            else {
                S.npar[a] = (uint8_t)np;
                for (uint32_t k = 0; k < np; k++) {
                    const uint32_t pn = par_node[po + k];
                    S.par_eid[a][k] = par_eid[po + k];
                    S.par_prev[a][k] = (idx_t)(n_prev ? mx_find<SH>(pk, pv, pn) : MX_NONE);
                    S.par_cur[a][k] = (idx_t)mx_find<SH>(ck, cv, pn);
                }
            }
            S.self_prev[a] = (idx_t)(n_prev ? mx_find<SH>(pk, pv, id) : MX_NONE);
            S.em_match[a] = emission[id] == xb;
        }
        __syncthreads();
        if (S.fail) break;
        // ---- 3. fm, fi (forward.rs:337-388)
        const XF mbp = S.mb[x], ibp = S.ib[x];
        const XF fb0 = xadd(xmul(mbp, lp.p_MM), xmul(ibp, lp.p_IM));
        const XF ib_cur = xmul(xadd(xmul(mbp, lp.p_MI), xmul(ibp, lp.p_II)), lp.p_random);
        for (uint32_t a = a_lo; a < n; a += MX_THREADS / MX_XB) {
            const uint32_t id = S.c_id[a];
            XF acc = xf_zero();
            const uint32_t np = S.npar[a];
            for (uint32_t k = 0; k < np; k++) {
                const uint32_t ps = S.par_prev[a][k];
                double pm = 0.0, pi = 0.0, pd = 0.0; int pe = 0;
                if (ps != MX_NONE) { pm = S.pm[ps][x]; pi = S.pi[ps][x]; pd = S.pd[ps][x]; pe = S.pe[ps][x]; }
                acc = xadd(acc, xf(trans_t[(size_t)S.par_eid[a][k] * n_x + xg] * (lp.p_MM * pm + lp.p_IM * pi + lp.p_DM * pd), pe));
            }
            acc = xadd(acc, xmul(fb0, init_t[(size_t)id * n_x + xg]));
            const XF m = xmul(acc, S.em_match[a] ? lp.p_match : lp.p_mismatch);
            const uint32_t sp = S.self_prev[a];
            double pm = 0.0, pi = 0.0, pd = 0.0; int pe = 0;
            if (sp != MX_NONE) { pm = S.pm[sp][x]; pi = S.pi[sp][x]; pd = S.pd[sp][x]; pe = S.pe[sp][x]; }
            const XF i = xf(lp.p_random * (lp.p_MI * pm + lp.p_II * pi + lp.p_DI * pd), pe);
            const int Em = xexp(m), Ei = xexp(i), Ec = Em > Ei ? Em : Ei;
            if (Ec == XF_ZERO_E) { S.cm[a][x] = 0.0; S.ci[a][x] = 0.0; S.cmie[a][x] = 0; }
            else { S.cm[a][x] = m.v == 0.0 ? 0.0 : m.v * pow2i(m.e - Ec); S.ci[a][x] = i.v == 0.0 ? 0.0 : i.v * pow2i(i.e - Ec); S.cmie[a][x] = Ec; }
        }
        __syncthreads();
        // ---- 4. fd0 + 4 x fdt over the same node list (forward.rs:423-466, non-adaptive)
        for (int t = 0; t < N_DEL_ROUNDS; t++) {
            for (uint32_t a = a_lo; a < n; a += MX_THREADS / MX_XB) {
                XF acc = xf_zero();
                const uint32_t np = S.npar[a];
                for (uint32_t k = 0; k < np; k++) {
                    const uint32_t pc = S.par_cur[a][k];
                    if (pc == MX_NONE) continue;
                    const double tr = trans_t[(size_t)S.par_eid[a][k] * n_x + xg];
                    if (t == 0) acc = xadd(acc, xf(tr * (lp.p_MD * S.cm[pc][x] + lp.p_ID * S.ci[pc][x]), S.cmie[pc][x]));
                    else acc = xadd(acc, xf(tr * lp.p_DD * S.dv[(t & 1) ^ 1][pc][x], S.de[(t & 1) ^ 1][pc][x]));
                }
                if (t == 0) acc = xadd(acc, xmul(ib_cur, lp.p_ID * init_t[(size_t)S.c_id[a] * n_x + xg]));
                S.dv[t & 1][a][x] = acc.v; S.de[t & 1][a][x] = acc.e;
                const XF tot = t == 0 ? acc : xadd(xf(S.cdv[a][x], S.cde[a][x]), acc);
                S.cdv[a][x] = tot.v; S.cde[a][x] = tot.e;
            }
            __syncthreads();
        }
        // ---- 5. fe (forward.rs:554-558), begin scalars ; pack the row into the "previous" arrays
        XF part = xf_zero();
        for (uint32_t a = a_lo; a < n; a += MX_THREADS / MX_XB) {
            const XF mi = xf(S.cm[a][x] + S.ci[a][x], S.cmie[a][x]), d = xf(S.cdv[a][x], S.cde[a][x]);
            part = xadd(part, xadd(mi, d));
            const Cell cl = cell_pack(xf(S.cm[a][x], S.cmie[a][x]), xf(S.ci[a][x], S.cmie[a][x]), d);
            S.pm[a][x] = cl.m; S.pi[a][x] = cl.i; S.pd[a][x] = cl.d; S.pe[a][x] = cl.e;
        }
        // lanes with the same candidate: l, l + XB, l + 2 XB, ... of every warp
        for (int o = MX_XB; o < 32; o <<= 1) {
            XF b; b.v = __shfl_xor_sync(0xffffffffu, part.v, o); b.e = __shfl_xor_sync(0xffffffffu, part.e, o);
            part = xadd(part, b);
        }
        if ((tid & 31) < MX_XB) S.esum[tid >> 5][x] = part;
        __syncthreads();
        if (tid < MX_XB) {
            XF e = xf_zero();
            for (int w = 0; w < MX_THREADS / 32; w++) e = xadd(e, S.esum[w][tid]);
            S.mb[tid] = xf_zero(); S.ib[tid] = xnorm(ib_cur);   // (this thread's candidate is tid)
            if (row + 1 == jb.len && (uint32_t)tid < jb.nx) out[jb.out0 + (size_t)tid * out_stride] = xnorm(xmul(e, lp.p_end));
        }
        n_prev = n; cur ^= 1; n_cells += n;
        (void)last;
        __syncthreads();
    }
    __syncthreads();
    if (tid == 0) { status[blockIdx.x] = S.fail; cells[blockIdx.x] = n_cells * jb.nx; }
}

__global__ void k_transpose_probs(const double* __restrict__ src, double* __restrict__ dst, uint32_t n, uint32_t n_x) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;   // dst[v][x] = src[x][v]
    if (i >= (size_t)n * n_x) return;
    const uint32_t v = (uint32_t)(i / n_x), x = (uint32_t)(i % n_x);
    dst[i] = src[(size_t)x * n + v];
}

// One launch of shape SH over `groups` (each with at most SH::XB candidates).  d_it / d_tt: init / trans transposed to [node][X].
template <class SH>
static int run_mapx_shape(dbgphmm_model* m, const std::vector<MapxGroup>& groups, const uint8_t* d_bases, const DevMappings& dmap, uint32_t out_stride,
                          const double* d_it, const double* d_tt, XF* h_final, std::vector<uint8_t>& failed, uint64_t* cells_out) {
    cudaStream_t st = MSET(m).stream;
    const uint32_t G = (uint32_t)groups.size(), X = m->n_batch;
    failed.assign(G, 0);
    if (G == 0) return DBGPHMM_OK;
    DevBuf b_jobs, b_out, b_status, b_cells;
    std::vector<MJob> mj(G);
    size_t n_out = 0;
    for (uint32_t g = 0; g < G; g++) {
        if (groups[g].nx > (uint32_t)SH::XB) { dbg_set_error("internal: mapx group wider than the kernel shape"); return DBGPHMM_ERR_INVALID; }
        mj[g].base_off = groups[g].base_off; mj[g].len = groups[g].len; mj[g].map_row0 = groups[g].map_row0;
        mj[g].x0 = groups[g].x0; mj[g].nx = groups[g].nx; mj[g].out0 = groups[g].out0;
        n_out = std::max<size_t>(n_out, (size_t)groups[g].out0 + (size_t)(groups[g].nx - 1) * out_stride + 1);
    }
    ST_TRY(dev_upload(b_jobs, mj, st));
    ST_TRY(b_out.alloc(sizeof(XF) * n_out)); ST_TRY(b_status.alloc(sizeof(int) * G)); ST_TRY(b_cells.alloc(sizeof(unsigned long long) * G));
    CUDA_TRY(cudaMemsetAsync(b_out.p, 0, sizeof(XF) * n_out, st));
    CUDA_TRY(cudaFuncSetAttribute(k_mapx<SH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(MXS<SH>)));
    k_mapx<SH><<<G, MX_THREADS, sizeof(MXS<SH>), st>>>(m->d_par_off, m->d_par_node, m->d_par_eid, m->d_emission, d_it, d_tt, X,
                                                       m->lin, b_jobs.as<MJob>(), d_bases, dmap.row_off, dmap.nodes, b_out.as<XF>(), out_stride,
                                                       b_status.as<int>(), b_cells.as<unsigned long long>());
    COUNT_LAUNCH();
    std::vector<int> status(G); std::vector<unsigned long long> cells(G);
    std::vector<XF> res(n_out);
    CUDA_TRY(cudaMemcpyAsync(status.data(), b_status.p, sizeof(int) * G, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(cells.data(), b_cells.p, sizeof(unsigned long long) * G, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(res.data(), b_out.p, sizeof(XF) * n_out, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    CUDA_TRY(cudaGetLastError());
    for (uint32_t g = 0; g < G; g++) {
        if (status[g]) { failed[g] = 1; continue; }
        for (uint32_t k = 0; k < groups[g].nx; k++) h_final[groups[g].out0 + (size_t)k * out_stride] = res[groups[g].out0 + (size_t)k * out_stride];
        if (cells_out) *cells_out += cells[g];
    }
    return DBGPHMM_OK;
}

// ln-space results are produced by the caller from out_final.  failed[g] != 0: group g must be re-run through k_sparse (a duplicated
// node in a row, a node with more than 8 parents, a row longer than the reference's 400-entry SparseVec).  Groups hold at most 8 candidates.
int run_mapx(dbgphmm_model* m, const std::vector<MapxGroup>& groups, const uint8_t* d_bases, const DevMappings& dmap, uint32_t out_stride,
             XF* h_final, std::vector<uint8_t>& failed, uint64_t* cells_out) {
    cudaStream_t st = MSET(m).stream;
    const uint32_t G = (uint32_t)groups.size(), X = m->n_batch;
    failed.assign(G, 0);
    if (G == 0) return DBGPHMM_OK;
    EvTimer tm(st, &g_times.sparse_ms);
    DevBuf b_it, b_tt;
    ST_TRY(b_it.alloc(sizeof(double) * (size_t)m->N * X)); ST_TRY(b_tt.alloc(sizeof(double) * (size_t)std::max<uint32_t>(m->E, 1) * X));
    {
        const size_t ni = (size_t)m->N * X, nt = (size_t)m->E * X;
        k_transpose_probs<<<(unsigned)((ni + 255) / 256), 256, 0, st>>>(m->d_init, b_it.as<double>(), m->N, X); COUNT_LAUNCH();
        if (nt) { k_transpose_probs<<<(unsigned)((nt + 255) / 256), 256, 0, st>>>(m->d_trans, b_tt.as<double>(), m->E, X); COUNT_LAUNCH(); }
    }
    ST_TRY(run_mapx_shape<MxCommon>(m, groups, d_bases, dmap, out_stride, b_it.as<double>(), b_tt.as<double>(), h_final, failed, cells_out));
    // what the common shape could not take (rows of more than 64 nodes, more than 4 parents): the wide shape, two candidates per CTA
    std::vector<MapxGroup> wide; std::vector<uint32_t> parent;
    for (uint32_t g = 0; g < G; g++) {
        if (!failed[g]) continue;
        for (uint32_t k = 0; k < groups[g].nx; k += MxWide::XB) {
            MapxGroup w = groups[g];
            w.x0 = groups[g].x0 + k; w.nx = std::min<uint32_t>(MxWide::XB, groups[g].nx - k); w.out0 = groups[g].out0 + k * out_stride;
            wide.push_back(w); parent.push_back(g);
        }
    }
    if (!wide.empty()) {
        if (getenv("DBGPHMM_TRACE")) fprintf(stderr, "[dbgphmm] mapx: %zu candidate groups through the wide shape\n", wide.size());
        std::vector<uint8_t> wfailed;
        ST_TRY(run_mapx_shape<MxWide>(m, wide, d_bases, dmap, out_stride, b_it.as<double>(), b_tt.as<double>(), h_final, wfailed, cells_out));
        for (uint32_t g = 0; g < G; g++) failed[g] = 0;
        for (size_t i = 0; i < wide.size(); i++) if (wfailed[i]) failed[parent[i]] = 1;   // (k_sparse then redoes the whole group)
    }
    return DBGPHMM_OK;
}
