// model.cu — device graph layout: relabelling, CSR adjacency, dense tiling plans, parameter sets.
//
// Replaces the petgraph DiGraph<PNode,PEdge> behind PHMMModel (hmmv2/common.rs:61-67,202-261) and the
// per-candidate rebuild of it in SeqGraph::to_phmm (graph/seq_graph.rs:160-223).
#include <algorithm>
#include <cmath>
#include <cstring>
#include <cstdio>
#include <cstdlib>
#include "model.h"
#include "dense.h"
#include "sparse.h"
#include "engine.h"

static thread_local std::string g_error;
std::atomic<unsigned long long> g_launch_count{0};
void dbg_set_error(const std::string& s) { g_error = s; }

extern "C" const char* dbgphmm_last_error(void) { return g_error.c_str(); }
extern "C" uint64_t dbgphmm_launch_count(int reset) {
    return reset ? g_launch_count.exchange(0) : g_launch_count.load();
}
extern "C" int dbgphmm_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    int ok = 0;
    for (int d = 0; d < n; d++) {
        cudaDeviceProp p;
        if (cudaGetDeviceProperties(&p, d) == cudaSuccess && p.major >= 10) ok++;
    }
    return ok;
}

// ------------------------------------------------------------------ params (hmmv2/params.rs:73-124)
extern "C" void dbgphmm_params_new(double p_mismatch, double p_gap_open, double p_gap_ext, double p_end,
                                   uint32_t n_active_nodes, uint32_t n_warmup, dbgphmm_params* q) {
    // Prob::from_prob(v) = ln v ; Prob::to_value() = exp(ln v)   (prob.rs:58-72)
    double l_mis = std::log(p_mismatch), l_open = std::log(p_gap_open), l_ext = std::log(p_gap_ext), l_end = std::log(p_end);
    q->p_mismatch = l_mis; q->p_gap_open = l_open; q->p_gap_ext = l_ext; q->p_end = l_end;
    q->p_DD = l_ext; q->p_II = l_ext; q->p_MI = l_open; q->p_MD = l_open; q->p_ID = l_open; q->p_DI = l_open;
    q->p_MM = std::log(1.0 - 2.0 * std::exp(l_open) - std::exp(l_end));
    q->p_DM = std::log(1.0 - std::exp(l_open) - std::exp(l_ext) - std::exp(l_end));
    q->p_IM = q->p_DM;
    q->p_match = std::log(1.0 - std::exp(l_mis));
    q->p_random = std::log(0.25);
    q->n_active_nodes = n_active_nodes;
    q->n_warmup = n_warmup;
    q->warmup_threshold = DBGPHMM_MAX_ACTIVE_NODES / 2;
    q->n_max_gaps = 4;
    q->active_node_max_ratio = 30.0;
}
extern "C" void dbgphmm_params_uniform(double p, dbgphmm_params* q) { dbgphmm_params_new(p, p, p, 0.00001, 40, 50, q); }

LinParams to_lin(const dbgphmm_params& p) {
    LinParams l;
    l.p_mismatch = std::exp(p.p_mismatch); l.p_match = std::exp(p.p_match); l.p_random = std::exp(p.p_random);
    l.p_end = std::exp(p.p_end);
    l.p_MM = std::exp(p.p_MM); l.p_IM = std::exp(p.p_IM); l.p_DM = std::exp(p.p_DM);
    l.p_MI = std::exp(p.p_MI); l.p_II = std::exp(p.p_II); l.p_DI = std::exp(p.p_DI);
    l.p_MD = std::exp(p.p_MD); l.p_ID = std::exp(p.p_ID); l.p_DD = std::exp(p.p_DD);
    l.n_active_nodes = p.n_active_nodes; l.n_warmup = p.n_warmup; l.warmup_threshold = p.warmup_threshold;
    l.n_max_gaps = p.n_max_gaps; l.active_node_max_ratio = p.active_node_max_ratio;
    return l;
}

template <class T>
static int upload(T** dptr, const std::vector<T>& h) {
    size_t bytes = std::max<size_t>(h.size(), 1) * sizeof(T);
    CUDA_TRY(cudaMalloc((void**)dptr, bytes));
    if (!h.empty()) CUDA_TRY(cudaMemcpy(*dptr, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
    return DBGPHMM_OK;
}

// ------------------------------------------------------------------ dense tiling plan
// up_off/up_node/up_eid: CSR of the upstream direction (parents for forward, children for backward).
// hops: depth of the upstream closure.  HALO_HOPS serves every single-row kernel ; 2 * HALO_HOPS is the register-stencil layout
// of the two-rows-per-launch kernel (dense.cu: k_dense_fwd2), whose tiles carry the dependency cone of two rows.
static int build_plan(DevPlan& P, uint32_t N, const std::vector<uint32_t>& up_off, const std::vector<uint32_t>& up_node,
                      const std::vector<uint32_t>& up_eid, int hops = HALO_HOPS) {
    const bool single = hops == HALO_HOPS;
    std::vector<uint32_t> chunk_start, loc_base, loc_node, nle, le_off, le_eid, fp_eid, fx_off, fx_eid;
    std::vector<uint16_t> le_idx, fp_idx, fx_idx;
    std::vector<uint32_t> rl_node, rl_eid, rx_off, rx_eid;
    std::vector<uint16_t> rl_par, rl_core, rx_idx;
    std::vector<uint8_t> rl_flag;
    std::vector<uint32_t> rpos(N, 0xffffffffu);  // node -> position in the register layout of the current tile
    std::vector<uint32_t> stamp(N, 0xffffffffu), lidx(N, 0);
    std::vector<uint32_t> local, depth_cnt(std::max(8, hops + 2));
    uint32_t start = 0, cidx = 0, max_local = 0;
    chunk_start.push_back(0);
    loc_base.push_back(0);
    std::vector<uint32_t> lay;   // register-stencil layout of the current tile: node per position, 0xffffffff = pad
    while (start < N) {
        uint32_t size = std::min<uint32_t>(DENSE_CORE, N - start);
        // Order the tile's nodes as [chain above the tile head | core | remaining halo chains] so that a node's first upstream
        // neighbour is usually the previous position, then pad so that every node whose first neighbour is elsewhere or
        // that has further upstream edges in the tile sits on slot 0 of a lane (position % DENSE_PER_LANE == 0): the kernel
        // specialises slot 0 and keeps the other slots a pure register stencil.  false = does not fit DENSE_LMAX.
        auto build_layout = [&]() -> bool {
            std::vector<uint32_t> order; order.reserve(DENSE_LMAX);
            std::vector<uint8_t> placed(local.size(), 0);
            auto in_tile = [&](uint32_t u) { return stamp[u] == cidx; };
            // core in ascending node order when the first upstream neighbour of v is usually v - 1 (forward: the relabelling
            // makes a parent the predecessor), descending when it is v + 1 (backward: upstream = children)
            uint32_t n_prev = 0, n_next = 0;
            for (uint32_t q = 0; q < size; q++) {
                const uint32_t x = local[q];
                if (up_off[x + 1] == up_off[x]) continue;
                const uint32_t u = up_node[up_off[x]];
                if (u + 1 == x) n_prev++; else if (u == x + 1) n_next++;
            }
            const bool descending = n_next > n_prev;
            std::vector<uint32_t> head;
            uint32_t v = descending ? local[size - 1] : local[0];
            for (int h = 0; h < hops; h++) {   // chain of first upstream neighbours above the tile head, deepest first
                if (up_off[v + 1] == up_off[v]) break;
                uint32_t u = up_node[up_off[v]];
                if (!in_tile(u) || lidx[u] < size || placed[lidx[u]]) break;
                placed[lidx[u]] = 1; head.push_back(u); v = u;
            }
            for (size_t q = head.size(); q-- > 0;) order.push_back(head[q]);
            for (uint32_t q = 0; q < size; q++) { order.push_back(local[descending ? size - 1 - q : q]); placed[q] = 1; }
            // remaining halo nodes: chains following "x is the first upstream neighbour of the next", deepest first
            for (size_t q = local.size(); q-- > size;) {   // local is depth-sorted: start from the deepest
                // The walk down from the top of local[q]'s chain may leave through another branch of a fork and never come back to
                // local[q] (its first neighbour can be shallower than itself when that one has a shorter way to the core): repeat
                // from local[q] until it is placed -- every pass places at least its top node, and stops climbing at placed nodes.
                while (!placed[q]) {
                    uint32_t x = local[q];
                    for (size_t steps = 0; steps < local.size(); steps++) {   // climb to the top of x's unplaced first-neighbour chain
                        if (up_off[x + 1] == up_off[x]) break;                // (bounded: a cycle of halo nodes has no top)
                        uint32_t u = up_node[up_off[x]];
                        if (!in_tile(u) || placed[lidx[u]]) break;
                        x = u;
                    }
                    for (;;) {   // walk down: append x, then an unplaced halo node whose first neighbour is x
                        placed[lidx[x]] = 1; order.push_back(x);
                        uint32_t nxt = 0xffffffffu;
                        for (size_t r = size; r < local.size(); r++) {
                            uint32_t y = local[r];
                            if (!placed[r] && up_off[y + 1] > up_off[y] && up_node[up_off[y]] == x) { nxt = y; break; }
                        }
                        if (nxt == 0xffffffffu) break;
                        x = nxt;
                    }
                }
            }
            lay.clear();
            for (size_t q = 0; q < order.size(); q++) {
                const uint32_t x = order[q];
                bool special = false, first = true;
                for (uint32_t a = up_off[x]; a < up_off[x + 1]; a++) {
                    const uint32_t u = up_node[a];
                    if (first) { first = false; if (in_tile(u) && !(q > 0 && order[q - 1] == u)) special = true; }
                    else if (in_tile(u)) special = true;
                }
                if (special) while (lay.size() % DENSE_PER_LANE) lay.push_back(0xffffffffu);
                lay.push_back(x);
                if (lay.size() > DENSE_LMAX) return false;
            }
            return true;
        };
        for (;;) {
            // closure of [start, start+size) over HALO_HOPS upstream hops
            local.clear();
            std::fill(depth_cnt.begin(), depth_cnt.end(), 0);
            for (uint32_t v = start; v < start + size; v++) { stamp[v] = cidx; lidx[v] = (uint32_t)local.size(); local.push_back(v); }
            depth_cnt[0] = size;
            size_t fb = 0, fe = local.size();
            bool ok = true;
            for (int h = 1; h <= hops && ok; h++) {
                for (size_t q = fb; q < fe; q++) {
                    uint32_t v = local[q];
                    for (uint32_t a = up_off[v]; a < up_off[v + 1]; a++) {
                        uint32_t u = up_node[a];
                        if (stamp[u] != cidx) {
                            stamp[u] = cidx; lidx[u] = (uint32_t)local.size(); local.push_back(u);
                            depth_cnt[h]++;
                        }
                    }
                    if (local.size() > DENSE_LMAX) { ok = false; break; }
                }
                fb = fe; fe = local.size();
            }
            if (ok && single) {  // the tile's edge list must fit as well (exact kernel)
                size_t n_edges_local = 0, n_need = 0;
                for (int h = 0; h < HALO_HOPS; h++) n_need += depth_cnt[h];
                for (size_t q = 0; q < n_need && q < local.size(); q++) n_edges_local += up_off[local[q] + 1] - up_off[local[q]];
                if (n_edges_local > DENSE_EMAX) ok = false;
                size_t n_extra = 0;
                for (size_t q = 0; q < n_need && q < local.size(); q++) { uint32_t dg = up_off[local[q] + 1] - up_off[local[q]]; if (dg > 1) n_extra += dg - 1; }
                if (n_extra > DENSE_XMAX) ok = false;
            }
            if (ok) ok = build_layout();
            if (ok) break;
            // undo stamps and retry with a smaller chunk
            for (uint32_t v : local) stamp[v] = 0xffffffffu;
            if (size == 1) { dbg_set_error("graph too dense: the 6-hop neighbourhood of one node exceeds the tile capacity"); return DBGPHMM_ERR_INVALID; }
            size = size > 16 ? size - 4 : std::max<uint32_t>(1, size / 2);
        }
        // emit chunk
        uint32_t base = (uint32_t)loc_node.size();
        uint32_t cum = 0;
        for (int h = 0; h < 8; h++) { cum += (h <= HALO_HOPS ? depth_cnt[h] : 0); nle.push_back(cum); }
        uint32_t n_need_edges = nle[nle.size() - 8 + (HALO_HOPS - 1)];  // depth <= 5 gather from upstream
        for (size_t j = 0; single && j < local.size(); j++) {
            uint32_t v = local[j];
            loc_node.push_back(v);
            le_off.push_back((uint32_t)le_idx.size());
            fx_off.push_back((uint32_t)fx_idx.size());
            uint16_t p0 = (uint16_t)j; uint32_t e0 = 0xffffffffu;
            if (j < n_need_edges) {
                for (uint32_t a = up_off[v]; a < up_off[v + 1]; a++) {
                    le_idx.push_back((uint16_t)lidx[up_node[a]]); le_eid.push_back(up_eid[a]);
                    if (a == up_off[v]) { p0 = (uint16_t)lidx[up_node[a]]; e0 = up_eid[a]; }
                    else { fx_idx.push_back((uint16_t)lidx[up_node[a]]); fx_eid.push_back(up_eid[a]); }
                }
            }
            fp_idx.push_back(p0); fp_eid.push_back(e0);
        }
        if (single) {
            le_off.push_back((uint32_t)le_idx.size());  // sentinel of this chunk
            fx_off.push_back((uint32_t)fx_idx.size());
        }
        // ---- register-stencil layout of this tile (positions computed by build_layout inside the sizing loop)
        {
            for (size_t q = 0; q < lay.size(); q++) if (lay[q] != 0xffffffffu) rpos[lay[q]] = (uint32_t)q;
            rl_core.push_back(0); rl_core.push_back((uint16_t)size);
            for (uint32_t q = 0; q < DENSE_LMAX; q++) {
                rx_off.push_back((uint32_t)rx_idx.size());
                const uint16_t prevpos = (uint16_t)(q > 0 ? q - 1 : 0);
                if (q >= lay.size() || lay[q] == 0xffffffffu) { rl_node.push_back(0xffffffffu); rl_par.push_back(prevpos); rl_eid.push_back(0xffffffffu); rl_flag.push_back(0); continue; }
                uint32_t x = lay[q];
                uint16_t p0 = prevpos; uint32_t e0 = 0xffffffffu; uint8_t fl = 0;
                bool first = true;
                for (uint32_t a = up_off[x]; a < up_off[x + 1]; a++) {
                    uint32_t u = up_node[a];
                    bool here = stamp[u] == cidx;
                    if (first) { first = false; if (here) { p0 = (uint16_t)rpos[u]; e0 = up_eid[a]; if (p0 != prevpos || q == 0) fl |= 1; } }
                    else if (here) { rx_idx.push_back((uint16_t)rpos[u]); rx_eid.push_back(up_eid[a]); fl |= 2; }
                }
                // the kernel reads the first neighbour of slot 0 of a lane from rl_par and of the other slots from the previous
                // register: build_layout put every node with fl != 0 on a slot 0
                if (fl && q % DENSE_PER_LANE) { dbg_set_error("internal: register layout misaligned"); return DBGPHMM_ERR_INVALID; }
                rl_node.push_back(x); rl_par.push_back(p0); rl_eid.push_back(e0); rl_flag.push_back(fl);
            }
            rx_off.push_back((uint32_t)rx_idx.size());
            {   // bit 2: the position is read by another one (slot-0 first neighbour, extra source): the kernel publishes only those
                const size_t fbase = rl_flag.size() - DENSE_LMAX;
                const size_t xbase = rx_off.size() - (DENSE_LMAX + 1);
                for (uint32_t q = 0; q < DENSE_LMAX; q++) {
                    if (q % DENSE_PER_LANE == 0) rl_flag[fbase + rl_par[fbase + q]] |= 4;
                    for (uint32_t e = rx_off[xbase + q]; e < rx_off[xbase + q + 1]; e++) rl_flag[fbase + rx_idx[e]] |= 4;
                }
            }
            for (uint32_t x : lay) if (x != 0xffffffffu) rpos[x] = 0xffffffffu;
        }
        max_local = std::max<uint32_t>(max_local, (uint32_t)local.size());
        for (uint32_t v : local) stamp[v] = 0xffffffffu;  // a node may be halo of several chunks
        start += size;
        cidx++;
        chunk_start.push_back(start);
        loc_base.push_back(base + (uint32_t)local.size());
        (void)base;
    }
    P.n_chunks = cidx;
    P.max_local = max_local;
    if (getenv("DBGPHMM_TRACE")) fprintf(stderr, "[dbgphmm] plan (%d hops): %u tiles for %u nodes (%.1f core nodes per tile, capacity %d)\n", hops, cidx, N, (double)N / cidx, DENSE_CORE);
    P.h_chunk_start = chunk_start;
    ST_TRY(upload(&P.chunk_start, chunk_start));
    ST_TRY(upload(&P.loc_base, loc_base));
    ST_TRY(upload(&P.loc_node, loc_node));
    ST_TRY(upload(&P.nle, nle));
    ST_TRY(upload(&P.le_off, le_off));
    ST_TRY(upload(&P.le_idx, le_idx));
    ST_TRY(upload(&P.le_eid, le_eid));
    ST_TRY(upload(&P.fp_idx, fp_idx)); ST_TRY(upload(&P.fp_eid, fp_eid));
    ST_TRY(upload(&P.fx_off, fx_off)); ST_TRY(upload(&P.fx_idx, fx_idx)); ST_TRY(upload(&P.fx_eid, fx_eid));
    ST_TRY(upload(&P.rl_node, rl_node)); ST_TRY(upload(&P.rl_par, rl_par)); ST_TRY(upload(&P.rl_eid, rl_eid)); ST_TRY(upload(&P.rl_flag, rl_flag));
    ST_TRY(upload(&P.rl_core, rl_core)); ST_TRY(upload(&P.rx_off, rx_off)); ST_TRY(upload(&P.rx_idx, rx_idx)); ST_TRY(upload(&P.rx_eid, rx_eid));
    return DBGPHMM_OK;
}

static void free_plan(DevPlan& P) {
    cudaFree(P.chunk_start); cudaFree(P.loc_base); cudaFree(P.loc_node); cudaFree(P.nle);
    cudaFree(P.le_off); cudaFree(P.le_idx); cudaFree(P.le_eid);
    cudaFree(P.fp_idx); cudaFree(P.fp_eid); cudaFree(P.fx_off); cudaFree(P.fx_idx); cudaFree(P.fx_eid);
    cudaFree(P.rl_node); cudaFree(P.rl_par); cudaFree(P.rl_eid); cudaFree(P.rl_flag); cudaFree(P.rl_core); cudaFree(P.rx_off); cudaFree(P.rx_idx); cudaFree(P.rx_eid);
    P = DevPlan();
}

// ------------------------------------------------------------------ graph build
int model_build_graph(dbgphmm_model* m, uint32_t N, uint32_t E, const uint32_t* src, const uint32_t* dst, const uint8_t* emission) {
    m->N = N; m->E = E;
    for (uint32_t e = 0; e < E; e++)
        if (src[e] >= N || dst[e] >= N) { dbg_set_error("edge endpoint out of range"); return DBGPHMM_ERR_INVALID; }
    m->e_src.assign(src, src + E); m->e_dst.assign(dst, dst + E);
    // adjacency in ORIGINAL ids, newest edge first (petgraph 0.6 linked lists)
    std::vector<uint32_t> ooff(N + 1, 0), ioff(N + 1, 0);
    for (uint32_t e = 0; e < E; e++) { ooff[src[e] + 1]++; ioff[dst[e] + 1]++; }
    for (uint32_t v = 0; v < N; v++) { ooff[v + 1] += ooff[v]; ioff[v + 1] += ioff[v]; }
    std::vector<uint32_t> oe(E), ie(E), ofill(ooff.begin(), ooff.end() - 1), ifill(ioff.begin(), ioff.end() - 1);
    for (uint32_t e = E; e-- > 0;) { oe[ofill[src[e]]++] = e; ie[ifill[dst[e]]++] = e; }
    auto indeg = [&](uint32_t v) { return ioff[v + 1] - ioff[v]; };
    auto outdeg = [&](uint32_t v) { return ooff[v + 1] - ooff[v]; };
    // ---- relabel: depth-first over maximal simple chains so that a node's parent is usually its predecessor.  A node with several
    // parents (the merge below a bubble) waits until all of them are labelled: the branches of a bubble then sit next to each other
    // in the new order, inside the same tile or two, instead of one branch following the main chain and the others being collected
    // far away at the end of the traversal -- where every one of them costs the tiles on both sides a 6- or 12-hop halo.
    std::vector<uint32_t> orig_of; orig_of.reserve(N);
    std::vector<uint8_t> visited(N, 0);
    std::vector<uint32_t> waiting(N);   // parents not labelled yet (self loops do not count)
    for (uint32_t v = 0; v < N; v++) { uint32_t w = 0; for (uint32_t a = ioff[v]; a < ioff[v + 1]; a++) w += src[ie[a]] != v; waiting[v] = w; }
    auto is_head = [&](uint32_t v) {
        if (indeg(v) != 1) return true;
        uint32_t p = src[ie[ioff[v]]];
        return outdeg(p) != 1 || p == v;
    };
    std::vector<uint32_t> stack, deferred;
    size_t deferred_head = 0;
    auto run_from = [&](uint32_t s) {
        stack.push_back(s);
        bool force = true;   // the start of a traversal, and a deferred node when nothing else is left (cycles), go in as they are
        for (;;) {
            if (stack.empty()) {
                while (deferred_head < deferred.size() && visited[deferred[deferred_head]]) deferred_head++;
                if (deferred_head == deferred.size()) break;
                stack.push_back(deferred[deferred_head++]); force = true;
            }
            uint32_t v = stack.back(); stack.pop_back();
            if (visited[v]) continue;
            if (!force && waiting[v] > 0) { deferred.push_back(v); continue; }
            force = false;
            // walk the chain
            for (;;) {
                visited[v] = 1; orig_of.push_back(v);
                for (uint32_t a = ooff[v]; a < ooff[v + 1]; a++) { uint32_t c = dst[oe[a]]; if (c != v && waiting[c] > 0) waiting[c]--; }
                if (outdeg(v) == 1) {
                    uint32_t c = dst[oe[ooff[v]]];
                    if (!visited[c] && !is_head(c)) { v = c; continue; }
                }
                break;
            }
            // children of the tail, pushed so that the newest-first first child is visited first
            for (uint32_t a = ooff[v + 1]; a-- > ooff[v];) { uint32_t c = dst[oe[a]]; if (!visited[c]) stack.push_back(c); }
        }
        deferred.clear(); deferred_head = 0;
    };
    for (uint32_t v = 0; v < N; v++) if (!visited[v] && indeg(v) == 0) run_from(v);
    for (uint32_t v = 0; v < N; v++) if (!visited[v] && is_head(v)) run_from(v);
    for (uint32_t v = 0; v < N; v++) if (!visited[v]) run_from(v);
    m->orig_of = orig_of;
    m->pos_of.assign(N, 0);
    for (uint32_t p = 0; p < N; p++) m->pos_of[orig_of[p]] = p;
    m->emission.resize(N);
    for (uint32_t p = 0; p < N; p++) m->emission[p] = emission[orig_of[p]];
    // ---- CSR in relabelled ids, per-node order = newest edge first
    m->par_off.assign(N + 1, 0); m->chi_off.assign(N + 1, 0);
    m->par_node.resize(E); m->par_eid.resize(E); m->chi_node.resize(E); m->chi_eid.resize(E);
    for (uint32_t p = 0; p < N; p++) {
        uint32_t v = orig_of[p];
        m->par_off[p + 1] = m->par_off[p] + indeg(v);
        m->chi_off[p + 1] = m->chi_off[p] + outdeg(v);
        for (uint32_t a = 0; a < indeg(v); a++) { uint32_t e = ie[ioff[v] + a]; m->par_node[m->par_off[p] + a] = m->pos_of[src[e]]; m->par_eid[m->par_off[p] + a] = e; }
        for (uint32_t a = 0; a < outdeg(v); a++) { uint32_t e = oe[ooff[v] + a]; m->chi_node[m->chi_off[p] + a] = m->pos_of[dst[e]]; m->chi_eid[m->chi_off[p] + a] = e; }
    }
    ST_TRY(upload(&m->d_pos_of, m->pos_of)); ST_TRY(upload(&m->d_orig_of, m->orig_of)); ST_TRY(upload(&m->d_emission, m->emission));
    ST_TRY(upload(&m->d_par_off, m->par_off)); ST_TRY(upload(&m->d_par_node, m->par_node)); ST_TRY(upload(&m->d_par_eid, m->par_eid));
    ST_TRY(upload(&m->d_chi_off, m->chi_off)); ST_TRY(upload(&m->d_chi_node, m->chi_node)); ST_TRY(upload(&m->d_chi_eid, m->chi_eid));
    {
        std::vector<uint4> pr(N), cr(N);
        for (uint32_t p = 0; p < N; p++) {
            const uint32_t po = m->par_off[p], pc = m->par_off[p + 1] - po, co = m->chi_off[p], cc = m->chi_off[p + 1] - co;
            m->max_deg = std::max(m->max_deg, std::max(pc, cc));
            pr[p] = make_uint4(po, pc, pc ? m->par_node[po] : 0u, pc ? m->par_eid[po] : 0u);
            cr[p] = make_uint4(co, cc, cc ? m->chi_node[co] : 0u, cc ? m->chi_eid[co] : 0u);
        }
        ST_TRY(upload(&m->d_par_rec, pr)); ST_TRY(upload(&m->d_chi_rec, cr));
    }
    ST_TRY(build_plan(m->fwd, N, m->par_off, m->par_node, m->par_eid));
    ST_TRY(build_plan(m->bwd, N, m->chi_off, m->chi_node, m->chi_eid));
    {   // layout of the two-rows-per-launch forward kernel ; a graph too branchy for 12-hop tiles simply goes without it
        if (build_plan(m->fwd2, N, m->par_off, m->par_node, m->par_eid, 2 * HALO_HOPS) != DBGPHMM_OK) free_plan(m->fwd2);
        if (build_plan(m->bwd2, N, m->chi_off, m->chi_node, m->chi_eid, 2 * HALO_HOPS) != DBGPHMM_OK) free_plan(m->bwd2);
    }
    return DBGPHMM_OK;
}

int model_upload_probs(dbgphmm_model* m, const double* log_init, const double* log_trans) {
    std::vector<double> init(m->N), trans(std::max<uint32_t>(m->E, 1));
    for (uint32_t p = 0; p < m->N; p++) init[p] = std::exp(log_init[m->orig_of[p]]);
    for (uint32_t e = 0; e < m->E; e++) trans[e] = std::exp(log_trans[e]);
    if (m->n_batch != 1 || !m->d_init) {
        cudaFree(m->d_init); cudaFree(m->d_trans); m->d_init = m->d_trans = nullptr;
        CUDA_TRY(cudaMalloc((void**)&m->d_init, sizeof(double) * std::max<uint32_t>(m->N, 1)));
        CUDA_TRY(cudaMalloc((void**)&m->d_trans, sizeof(double) * std::max<uint32_t>(m->E, 1)));
        m->n_batch = 1;
    }
    CUDA_TRY(cudaMemcpy(m->d_init, init.data(), sizeof(double) * m->N, cudaMemcpyHostToDevice));
    if (m->E) CUDA_TRY(cudaMemcpy(m->d_trans, trans.data(), sizeof(double) * m->E, cudaMemcpyHostToDevice));
    return DBGPHMM_OK;
}

// tiles that intersect the closure of each tile's nodes over `hops` upstream hops (up_* = parents: forward tiles ; children: backward)
static int build_roi(const std::vector<uint32_t>& cs, uint32_t T, uint32_t N, uint32_t hops, const std::vector<uint32_t>& up_off,
                     const std::vector<uint32_t>& up_node, uint32_t** d_off, uint32_t** d_tile, uint32_t** d_tile_of) {
    std::vector<uint32_t> tile_of(N);
    for (uint32_t t = 0; t < T; t++) for (uint32_t v = cs[t]; v < cs[t + 1]; v++) tile_of[v] = t;
    std::vector<uint32_t> roi_off(1, 0), roi_tile, stamp(N, 0xffffffffu), tstamp(T, 0xffffffffu), frontier, next;
    for (uint32_t t = 0; t < T; t++) {
        frontier.clear();
        for (uint32_t v = cs[t]; v < cs[t + 1]; v++) { stamp[v] = t; frontier.push_back(v); }
        tstamp[t] = t; roi_tile.push_back(t);
        for (uint32_t h = 0; h < hops && !frontier.empty(); h++) {
            next.clear();
            for (uint32_t v : frontier)
                for (uint32_t a = up_off[v]; a < up_off[v + 1]; a++) {
                    uint32_t u = up_node[a];
                    if (stamp[u] != t) {
                        stamp[u] = t; next.push_back(u);
                        uint32_t tu = tile_of[u];
                        if (tstamp[tu] != t) { tstamp[tu] = t; roi_tile.push_back(tu); }
                    }
                }
            frontier.swap(next);
        }
        roi_off.push_back((uint32_t)roi_tile.size());
    }
    ST_TRY(upload(d_off, roi_off)); ST_TRY(upload(d_tile, roi_tile)); ST_TRY(upload(d_tile_of, tile_of));
    return DBGPHMM_OK;
}

int model_ensure_roi(dbgphmm_model* m) {
    const uint32_t W = m->params.n_warmup;
    if (m->d_roi_off && m->roi_warmup == W) return DBGPHMM_OK;
    cudaFree(m->d_roi_off); cudaFree(m->d_roi_tile); cudaFree(m->d_tile_of);
    cudaFree(m->d_roi_off_b); cudaFree(m->d_roi_tile_b); cudaFree(m->d_tile_of_b);
    m->d_roi_off = m->d_roi_tile = m->d_tile_of = m->d_roi_off_b = m->d_roi_tile_b = m->d_tile_of_b = nullptr;
    ST_TRY(build_roi(m->fwd.h_chunk_start, m->fwd.n_chunks, m->N, HALO_HOPS * W, m->par_off, m->par_node, &m->d_roi_off, &m->d_roi_tile, &m->d_tile_of));
    ST_TRY(build_roi(m->bwd.h_chunk_start, m->bwd.n_chunks, m->N, HALO_HOPS * W, m->chi_off, m->chi_node, &m->d_roi_off_b, &m->d_roi_tile_b, &m->d_tile_of_b));
    m->roi_warmup = W;
    return DBGPHMM_OK;
}

void model_free(dbgphmm_model* m) {
    if (!m) return;
    cudaSetDevice(m->device);
    cudaFree(m->d_roi_off); cudaFree(m->d_roi_tile); cudaFree(m->d_tile_of);
    cudaFree(m->d_roi_off_b); cudaFree(m->d_roi_tile_b); cudaFree(m->d_tile_of_b);
    cudaFree(m->d_pos_of); cudaFree(m->d_orig_of); cudaFree(m->d_emission);
    cudaFree(m->d_par_off); cudaFree(m->d_par_node); cudaFree(m->d_par_eid);
    cudaFree(m->d_chi_off); cudaFree(m->d_chi_node); cudaFree(m->d_chi_eid);
    cudaFree(m->d_par_rec); cudaFree(m->d_chi_rec);
    cudaFree(m->d_init); cudaFree(m->d_trans);
    free_plan(m->fwd); free_plan(m->bwd); free_plan(m->fwd2); free_plan(m->bwd2);
    cache_set_stream(nullptr);   // (this handle's streams are about to go)
    cache_trim();
    for (auto& ss : m->ss) {
        cudaFree(ss.d_jstep);
        if (ss.stream) cudaStreamDestroy(ss.stream);
        if (ss.aux) cudaStreamDestroy(ss.aux);
        if (ss.ev_fork) cudaEventDestroy(ss.ev_fork);
        if (ss.ev_join) cudaEventDestroy(ss.ev_join);
    }
    delete m;
}

thread_local int tl_stream_set = 0;

// ------------------------------------------------------------------ copy numbers -> parameters on the device
// graph/seq_graph.rs:110-135,160-273 with edge copy numbers None (multi_dbg.rs:1386-1387,1434-1437).
__global__ void k_copy_total(const uint32_t* __restrict__ cn, const uint8_t* __restrict__ emission, const uint32_t* __restrict__ orig_of,
                             uint32_t N, int mode, unsigned long long* __restrict__ total) {
    uint32_t x = blockIdx.y;
    unsigned long long s = 0;
    for (uint32_t p = blockIdx.x * blockDim.x + threadIdx.x; p < N; p += gridDim.x * blockDim.x) {
        if (emission[p] != 'n') {
            unsigned long long c = cn[(size_t)x * N + orig_of[p]];
            if (mode == 1 && c < 1) c = 1;
            if (mode == 2) c = 1;
            s += c;
        }
    }
    for (int o = 16; o; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0 && s) atomicAdd(&total[x], s);
}
__global__ void k_copy_to_probs(const uint32_t* __restrict__ cn, const uint8_t* __restrict__ emission, const uint32_t* __restrict__ orig_of,
                                const uint32_t* __restrict__ chi_off, const uint32_t* __restrict__ chi_node, const uint32_t* __restrict__ chi_eid,
                                uint32_t N, uint32_t E, int mode, const unsigned long long* __restrict__ total,
                                double* __restrict__ init, double* __restrict__ trans) {
    uint32_t x = blockIdx.y;
    uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= N) return;
    auto eff = [&](uint32_t q) -> unsigned long long {
        if (emission[q] == 'n') return 0ull;
        unsigned long long c = cn[(size_t)x * N + orig_of[q]];
        if (mode == 1 && c < 1) c = 1;
        if (mode == 2) c = 1;
        return c;
    };
    unsigned long long c = eff(p);
    init[(size_t)x * N + p] = (emission[p] != 'n' && total[x] > 0) ? (double)c / (double)total[x] : 0.0;
    unsigned long long ct = 0;
    for (uint32_t a = chi_off[p]; a < chi_off[p + 1]; a++) ct += eff(chi_node[a]);
    for (uint32_t a = chi_off[p]; a < chi_off[p + 1]; a++) {
        uint32_t q = chi_node[a];
        unsigned long long cq = eff(q);
        trans[(size_t)x * E + chi_eid[a]] = (emission[q] != 'n' && ct > 0) ? (double)cq / (double)ct : 0.0;
    }
}

extern "C" int dbgphmm_model_set_copy_nums_batch(dbgphmm_model* m, uint32_t n_batch, const uint32_t* copy_nums, int mode) try {
    if (!m || !copy_nums || n_batch == 0 || mode < 0 || mode > 2) { dbg_set_error("set_copy_nums_batch: bad argument"); return DBGPHMM_ERR_INVALID; }
    CUDA_TRY(cudaSetDevice(m->device));
    uint32_t N = m->N, E = m->E;
    uint32_t* d_cn = nullptr; unsigned long long* d_total = nullptr;
    CUDA_TRY(cudaMalloc((void**)&d_cn, sizeof(uint32_t) * (size_t)n_batch * N));
    CUDA_TRY(cudaMalloc((void**)&d_total, sizeof(unsigned long long) * n_batch));
    CUDA_TRY(cudaMemcpyAsync(d_cn, copy_nums, sizeof(uint32_t) * (size_t)n_batch * N, cudaMemcpyHostToDevice, MSET(m).stream));
    CUDA_TRY(cudaMemsetAsync(d_total, 0, sizeof(unsigned long long) * n_batch, MSET(m).stream));
    cudaFree(m->d_init); cudaFree(m->d_trans); m->d_init = m->d_trans = nullptr;
    CUDA_TRY(cudaMalloc((void**)&m->d_init, sizeof(double) * (size_t)n_batch * std::max<uint32_t>(N, 1)));
    CUDA_TRY(cudaMalloc((void**)&m->d_trans, sizeof(double) * (size_t)n_batch * std::max<uint32_t>(E, 1)));
    m->n_batch = n_batch;
    dim3 g1(std::min<uint32_t>((N + 255) / 256, 1024), n_batch), g2((N + 255) / 256, n_batch);
    k_copy_total<<<g1, 256, 0, MSET(m).stream>>>(d_cn, m->d_emission, m->d_orig_of, N, mode, d_total); COUNT_LAUNCH();
    k_copy_to_probs<<<g2, 256, 0, MSET(m).stream>>>(d_cn, m->d_emission, m->d_orig_of, m->d_chi_off, m->d_chi_node, m->d_chi_eid, N, E, mode,
                                                d_total, m->d_init, m->d_trans); COUNT_LAUNCH();
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaStreamSynchronize(MSET(m).stream));
    cudaFree(d_cn); cudaFree(d_total);
    return DBGPHMM_OK;
} ABI_CATCH

extern "C" int dbgphmm_model_get_probs(const dbgphmm_model* m, uint32_t x, double* log_init, double* log_trans) try {
    if (!m || x >= m->n_batch) { dbg_set_error("get_probs: bad argument"); return DBGPHMM_ERR_INVALID; }
    CUDA_TRY(cudaSetDevice(m->device));
    std::vector<double> init(m->N), trans(m->E);
    CUDA_TRY(cudaMemcpy(init.data(), m->d_init + (size_t)x * m->N, sizeof(double) * m->N, cudaMemcpyDeviceToHost));
    if (m->E) CUDA_TRY(cudaMemcpy(trans.data(), m->d_trans + (size_t)x * m->E, sizeof(double) * m->E, cudaMemcpyDeviceToHost));
    for (uint32_t p = 0; p < m->N; p++) log_init[m->orig_of[p]] = std::log(init[p]);
    for (uint32_t e = 0; e < m->E; e++) log_trans[e] = std::log(trans[e]);
    return DBGPHMM_OK;
} ABI_CATCH

extern "C" int dbgphmm_model_create(uint32_t n_nodes, uint32_t n_edges, const uint32_t* edge_src, const uint32_t* edge_dst,
                                    const uint8_t* emission, const double* log_init, const double* log_trans,
                                    const dbgphmm_params* params, int device, uint64_t mem_budget_bytes, dbgphmm_model** out) try {
    if (!out || !params || !emission || !log_init || (n_edges && (!edge_src || !edge_dst || !log_trans)) || n_nodes == 0) {
        dbg_set_error("model_create: bad argument"); return DBGPHMM_ERR_INVALID;
    }
    if (params->n_max_gaps != 4) { dbg_set_error("model_create: n_max_gaps must be 4 (table.rs:17)"); return DBGPHMM_ERR_INVALID; }
    if (params->n_active_nodes == 0 || params->n_active_nodes >= DBGPHMM_MAX_ACTIVE_NODES || params->n_warmup == 0) {
        dbg_set_error("model_create: need 0 < n_active_nodes < 400 and n_warmup > 0 (params.rs:81-83)"); return DBGPHMM_ERR_INVALID;
    }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) {
        cudaGetLastError();
        dbg_set_error("model_create: no CUDA device (this library has no CPU path)"); return DBGPHMM_ERR_CUDA;
    }
    CUDA_TRY(cudaSetDevice(device));
    dbgphmm_model* m = new dbgphmm_model();
    static std::atomic<uint64_t> next_serial{0};
    m->serial = ++next_serial;
    m->device = device; m->params = *params; m->lin = to_lin(*params);
    int st = DBGPHMM_OK;
    for (auto& ss : m->ss)
        if (cudaStreamCreateWithFlags(&ss.stream, cudaStreamNonBlocking) != cudaSuccess || cudaStreamCreateWithFlags(&ss.aux, cudaStreamNonBlocking) != cudaSuccess ||
            cudaEventCreateWithFlags(&ss.ev_fork, cudaEventDisableTiming) != cudaSuccess || cudaEventCreateWithFlags(&ss.ev_join, cudaEventDisableTiming) != cudaSuccess) {
            dbg_set_error("stream create failed"); st = DBGPHMM_ERR_CUDA;
        }
    if (st == DBGPHMM_OK) st = model_build_graph(m, n_nodes, n_edges, edge_src, edge_dst, emission);
    if (st == DBGPHMM_OK) st = model_upload_probs(m, log_init, log_trans);
    if (st == DBGPHMM_OK) {
        size_t fr = 0, tot = 0;
        cudaMemGetInfo(&fr, &tot);
        m->mem_budget = mem_budget_bytes ? mem_budget_bytes : (uint64_t)(fr * 0.88);
        m->mem_budget_fixed = mem_budget_bytes != 0;
        st = dense_configure(m);
        if (st == DBGPHMM_OK) st = sparse_configure(m);
    }
    if (st != DBGPHMM_OK) { model_free(m); return st; }
    *out = m;
    return DBGPHMM_OK;
} ABI_CATCH
extern "C" void dbgphmm_model_destroy(dbgphmm_model* m) { model_free(m); }
extern "C" int dbgphmm_model_set_params(dbgphmm_model* m, const dbgphmm_params* p) try {
    if (!m || !p || p->n_max_gaps != 4) { dbg_set_error("set_params: bad argument"); return DBGPHMM_ERR_INVALID; }
    m->params = *p; m->lin = to_lin(*p);
    return DBGPHMM_OK;
} ABI_CATCH
extern "C" int dbgphmm_model_set_probs(dbgphmm_model* m, const double* log_init, const double* log_trans) try {
    if (!m || !log_init) { dbg_set_error("set_probs: bad argument"); return DBGPHMM_ERR_INVALID; }
    CUDA_TRY(cudaSetDevice(m->device));
    return model_upload_probs(m, log_init, log_trans);
} ABI_CATCH
extern "C" uint32_t dbgphmm_model_n_nodes(const dbgphmm_model* m) { return m ? m->N : 0; }
extern "C" uint32_t dbgphmm_model_n_batch(const dbgphmm_model* m) { return m ? m->n_batch : 0; }
