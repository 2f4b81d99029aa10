// Sanitizer run of the library's HOST logic without a GPU (tests/test_host_logic_asan.py builds csrc/model.cu and csrc/engine.cu
// with -fsanitize=address,undefined and links them with tests/native/cuda_stub.cpp instead of the CUDA runtime):
//   plan <graph.bin>...   dbgphmm_model_create on each graph (relabelling, adjacency, the four tiling plans, the recompute cones),
//                         then every table a dense kernel indexes with is checked against the graph: positions in range, the
//                         first-neighbour / extra-edge tables name real edges of the right direction, slot alignment, source flags,
//                         every upstream edge of a core node present in its tile
//   cache                 the stream-ordered block cache of engine.cu: reuse rules, ownership, events, trimming, budget
// graph.bin = u32 N, u32 E, u32 src[E], u32 dst[E], u8 emission[N], f64 log_init[N], f64 log_trans[E], then the bytes of dbgphmm_params.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <set>
#include <string>
#include <vector>
#include "../../dbgphmm_b200/csrc/engine.h"

extern size_t stub_device_bytes, stub_allocs, stub_frees, stub_stream_waits, stub_device_syncs;
size_t stub_bytes_in_use();

// ---- the device side of the library is not part of this program
int dense_configure(dbgphmm_model*) { return DBGPHMM_OK; }
int sparse_configure(dbgphmm_model*) { return DBGPHMM_OK; }
#define NOT_HERE { fprintf(stderr, "device path called in the host-logic program\n"); abort(); }
int sparse_run(dbgphmm_model*, const SJob*, uint32_t, const SparseIO&, uint32_t, int, uint32_t) NOT_HERE
int dense_select(dbgphmm_model*, const DensePool&, const SelectReq*, uint32_t, const int*, uint32_t*, uint32_t*) NOT_HERE
int step_products(dbgphmm_model*, const StepProducts&, const DensePool&, const DJob*, uint32_t, uint32_t, int, uint32_t) NOT_HERE
bool dense_can_pair(const dbgphmm_model*) NOT_HERE
uint32_t dense_pair_tiles(const dbgphmm_model*, int) NOT_HERE
uint32_t sparse_gather_cap(const dbgphmm_model*, uint32_t) NOT_HERE
uint32_t sparse_rescue_cap(uint32_t) NOT_HERE
uint32_t sparse_default_cap() NOT_HERE
int dense_forward_pair(dbgphmm_model*, const DensePool&, const DJob*, uint32_t, uint32_t, const uint8_t*, RowDesc*, XF*, int*, size_t, bool) NOT_HERE
int dense_backward_pair(dbgphmm_model*, const DensePool&, const DJob*, uint32_t, uint32_t, const uint8_t*, RowDesc*, XF*, int*, size_t, bool) NOT_HERE
int dense_forward_step(dbgphmm_model*, const DensePool&, const DJob*, uint32_t, uint32_t, const uint8_t*, RowDesc*, const int*, XF*, unsigned long long*, size_t) NOT_HERE
int dense_backward_step(dbgphmm_model*, const DensePool&, const DJob*, uint32_t, uint32_t, const uint8_t*, RowDesc*, const int*, XF*, unsigned long long*, size_t) NOT_HERE
int sparse_gather_prev0(dbgphmm_model*, int, uint32_t, uint32_t, const uint32_t*, const uint32_t*, const uint64_t*, const char*, uint64_t, uint32_t, uint32_t, char*, uint32_t*, int*) NOT_HERE
int dense_forward_step_list(dbgphmm_model*, const DensePool&, const DJob*, uint32_t, const uint8_t*, const RowDesc*, XF*, const unsigned long long*) NOT_HERE
int dense_backward_step_list(dbgphmm_model*, const DensePool&, const DJob*, uint32_t, const uint8_t*, XF*, const unsigned long long*) NOT_HERE

#define CHECK(cond, ...) do { if (!(cond)) { fprintf(stderr, "CHECK failed %s:%d: %s -- ", __FILE__, __LINE__, #cond); fprintf(stderr, __VA_ARGS__); fprintf(stderr, "\n"); exit(1); } } while (0)

struct Stats { uint64_t tiles = 0, positions = 0, core = 0, pads = 0, plain = 0, extras = 0; };

// up = parents (forward plans) or children (backward plans), CSR over relabelled ids
static Stats check_plan(const dbgphmm_model* m, const DevPlan& P, bool fwd, const char* name, int hops) {
    Stats st;
    const uint32_t N = m->N, L = DENSE_LMAX, PL = DENSE_PER_LANE;
    const std::vector<uint32_t>& up_off = fwd ? m->par_off : m->chi_off;
    const std::vector<uint32_t>& up_node = fwd ? m->par_node : m->chi_node;
    const std::vector<uint32_t>& up_eid = fwd ? m->par_eid : m->chi_eid;
    CHECK(P.n_chunks > 0 && P.chunk_start[0] == 0 && P.chunk_start[P.n_chunks] == N, "%s: chunk bounds", name);
    std::vector<int> pos_of(N, -1);
    auto edge_ok = [&](uint32_t eid, uint32_t up, uint32_t x) {   // edge eid leads from the upstream node `up` to x, in original ids
        if (eid >= m->E) return false;
        const uint32_t a = m->orig_of[up], b = m->orig_of[x];
        return fwd ? (m->e_src[eid] == a && m->e_dst[eid] == b) : (m->e_src[eid] == b && m->e_dst[eid] == a);
    };
    for (uint32_t c = 0; c < P.n_chunks; c++) {
        const uint32_t g0 = P.chunk_start[c], g1 = P.chunk_start[c + 1];
        CHECK(g0 < g1 && g1 <= N && g1 - g0 <= DENSE_CORE, "%s tile %u: core [%u, %u)", name, c, g0, g1);
        const uint32_t* node = P.rl_node + (size_t)c * L; const uint16_t* par = P.rl_par + (size_t)c * L;
        const uint32_t* eid = P.rl_eid + (size_t)c * L; const uint8_t* flag = P.rl_flag + (size_t)c * L;
        const uint32_t* xo = P.rx_off + (size_t)c * (L + 1);
        std::vector<uint32_t> here;
        uint32_t n_core = 0; bool plain = true;
        for (uint32_t q = 0; q < L; q++) {
            CHECK(xo[q] <= xo[q + 1], "%s tile %u: rx_off not monotone at %u", name, c, q);
            if (node[q] == 0xffffffffu) { st.pads++; CHECK(xo[q] == xo[q + 1] && par[q] < L, "%s tile %u: pad %u carries edges", name, c, q); continue; }
            CHECK(node[q] < N && pos_of[node[q]] < 0, "%s tile %u: position %u node %u out of range or placed twice", name, c, q, node[q]);
            pos_of[node[q]] = (int)q; here.push_back(node[q]);
            if (node[q] >= g0 && node[q] < g1) n_core++;
        }
        CHECK(n_core == g1 - g0, "%s tile %u: %u of %u core nodes placed", name, c, n_core, g1 - g0);
        std::vector<uint8_t> is_source(L, 0);
        for (uint32_t q = 0; q < L; q++) {
            if (q % PL == 0) { CHECK(par[q] < L, "%s tile %u: rl_par[%u] = %u", name, c, q, par[q]); is_source[par[q]] = 1; }   // (what slot 0 of a lane reads)
            if (node[q] == 0xffffffffu) continue;
            const uint32_t x = node[q];
            const uint32_t deg = up_off[x + 1] - up_off[x];
            // first upstream neighbour
            uint32_t n_in_tile = 0;
            if (eid[q] != 0xffffffffu) {
                CHECK(deg > 0 && par[q] < L && node[par[q]] != 0xffffffffu, "%s tile %u pos %u: first neighbour position %u", name, c, q, par[q]);
                CHECK(node[par[q]] == up_node[up_off[x]] && eid[q] == up_eid[up_off[x]], "%s tile %u pos %u: not the first upstream neighbour in CSR order", name, c, q);
                CHECK(edge_ok(eid[q], node[par[q]], x), "%s tile %u pos %u: edge %u does not join the two nodes", name, c, q, eid[q]);
                CHECK(((flag[q] & 1) != 0) == (par[q] + 1 != q || q == 0), "%s tile %u pos %u: flag bit 0", name, c, q);
                n_in_tile++;
            } else {
                CHECK(!(flag[q] & 1), "%s tile %u pos %u: flag bit 0 without an edge", name, c, q);
                CHECK(deg == 0 || pos_of[up_node[up_off[x]]] < 0, "%s tile %u pos %u: first neighbour is in the tile but not linked", name, c, q);
            }
            // the other upstream neighbours present in the tile, in CSR order
            uint32_t e = xo[q];
            for (uint32_t a = up_off[x] + 1; a < up_off[x + 1]; a++) {
                const uint32_t u = up_node[a];
                if (pos_of[u] < 0) continue;
                CHECK(e < xo[q + 1] && P.rx_idx[e] == (uint32_t)pos_of[u] && P.rx_eid[e] == up_eid[a] && edge_ok(P.rx_eid[e], u, x), "%s tile %u pos %u: extra edge list", name, c, q);
                is_source[P.rx_idx[e]] = 1; e++; n_in_tile++; st.extras++;
            }
            CHECK(e == xo[q + 1], "%s tile %u pos %u: %u surplus extra edges", name, c, q, xo[q + 1] - e);
            CHECK(((flag[q] & 2) != 0) == (xo[q + 1] > xo[q]), "%s tile %u pos %u: flag bit 1", name, c, q);
            if (q % PL) CHECK(!(flag[q] & 3) && (eid[q] == 0xffffffffu || par[q] + 1 == q), "%s tile %u pos %u: a special node off slot 0", name, c, q);
            if (x >= g0 && x < g1) CHECK(n_in_tile == deg, "%s tile %u pos %u: core node %u has %u of %u upstream edges in the tile", name, c, q, x, n_in_tile, deg);
        }
        {   // the tile holds the whole upstream closure of its core over `hops` hops: a node nearer than that has every upstream edge here
            std::vector<uint32_t> frontier, next; std::vector<int> depth(L, -1);
            for (uint32_t x : here) if (x >= g0 && x < g1) { depth[pos_of[x]] = 0; frontier.push_back(x); }
            for (int h = 0; h < hops; h++) {
                next.clear();
                for (uint32_t x : frontier)
                    for (uint32_t a = up_off[x]; a < up_off[x + 1]; a++) {
                        const uint32_t u = up_node[a];
                        CHECK(pos_of[u] >= 0, "%s tile %u: node %u at depth %d lacks its upstream neighbour %u", name, c, x, h, u);
                        if (depth[pos_of[u]] < 0) { depth[pos_of[u]] = h + 1; next.push_back(u); }
                    }
                frontier.swap(next);
            }
            for (uint32_t x : here) CHECK(depth[pos_of[x]] >= 0, "%s tile %u: node %u is not within %d upstream hops of the core", name, c, x, hops);
            // and every edge between two nodes of the tile whose head is nearer than `hops` is in the tables (first neighbour or extra)
            for (uint32_t x : here) {
                if (depth[pos_of[x]] >= hops) continue;
                const uint32_t q = (uint32_t)pos_of[x], deg = up_off[x + 1] - up_off[x];
                CHECK((eid[q] != 0xffffffffu ? 1u : 0u) + (xo[q + 1] - xo[q]) == deg, "%s tile %u pos %u: %u upstream edges, %u in the tables", name, c, q, deg, (eid[q] != 0xffffffffu ? 1u : 0u) + (xo[q + 1] - xo[q]));
            }
        }
        for (uint32_t q = 0; q < L; q++) {
            if (is_source[q]) CHECK(flag[q] & 4, "%s tile %u pos %u: read by another position but not flagged as a source", name, c, q);
            if ((flag[q] & 4) && q % PL != PL - 1) plain = false;
        }
        for (uint32_t x : here) pos_of[x] = -1;
        st.tiles++; st.positions += here.size(); st.core += g1 - g0; st.plain += plain;
    }
    return st;
}

static std::vector<char> slurp(const char* path) {
    FILE* f = fopen(path, "rb");
    if (!f) { fprintf(stderr, "cannot open %s\n", path); exit(2); }
    std::vector<char> b; char buf[1 << 16]; size_t n;
    while ((n = fread(buf, 1, sizeof buf, f)) > 0) b.insert(b.end(), buf, buf + n);
    fclose(f);
    return b;
}

static int run_plan(const char* path) {
    const std::vector<char> b = slurp(path);
    const char* p = b.data();
    uint32_t N, E; memcpy(&N, p, 4); memcpy(&E, p + 4, 4); p += 8;
    std::vector<uint32_t> src(E), dst(E); std::vector<uint8_t> em(N); std::vector<double> li(N), lt(E);
    auto take = [&](void* to, size_t n) { if (n) memcpy(to, p, n); p += n; };
    take(src.data(), 4ull * E); take(dst.data(), 4ull * E); take(em.data(), N); take(li.data(), 8ull * N); take(lt.data(), 8ull * E);
    dbgphmm_params par; CHECK((size_t)(b.data() + b.size() - p) == sizeof par, "%s: size", path); memcpy(&par, p, sizeof par);
    dbgphmm_model* m = nullptr;
    const int st = dbgphmm_model_create(N, E, src.data(), dst.data(), em.data(), li.data(), lt.data(), &par, 0, 0, &m);
    if (st != DBGPHMM_OK) { printf("%s: rejected (%s)\n", path, dbgphmm_last_error()); return 0; }
    // relabelling is a permutation, adjacency mirrors the edge list
    std::vector<uint8_t> seen(N, 0);
    for (uint32_t q = 0; q < N; q++) { CHECK(m->orig_of[q] < N && !seen[m->orig_of[q]] && m->pos_of[m->orig_of[q]] == q, "relabelling"); seen[m->orig_of[q]] = 1; }
    CHECK(m->par_off[N] == E && m->chi_off[N] == E, "adjacency sizes");
    for (uint32_t q = 0; q < N; q++) {
        for (uint32_t a = m->par_off[q]; a < m->par_off[q + 1]; a++) CHECK(m->e_dst[m->par_eid[a]] == m->orig_of[q] && m->e_src[m->par_eid[a]] == m->orig_of[m->par_node[a]], "parents CSR");
        for (uint32_t a = m->chi_off[q]; a < m->chi_off[q + 1]; a++) CHECK(m->e_src[m->chi_eid[a]] == m->orig_of[q] && m->e_dst[m->chi_eid[a]] == m->orig_of[m->chi_node[a]], "children CSR");
        for (uint32_t a = m->par_off[q] + 1; a < m->par_off[q + 1]; a++) CHECK(m->par_eid[a - 1] > m->par_eid[a], "parents not newest edge first");
        for (uint32_t a = m->chi_off[q] + 1; a < m->chi_off[q + 1]; a++) CHECK(m->chi_eid[a - 1] > m->chi_eid[a], "children not newest edge first");
    }
    printf("%s: N=%u E=%u max_deg=%u\n", path, N, E, m->max_deg);
    struct { const DevPlan* P; bool fwd; const char* name; int hops; } plans[4] = {{&m->fwd, true, "fwd", HALO_HOPS}, {&m->bwd, false, "bwd", HALO_HOPS}, {&m->fwd2, true, "fwd2", 2 * HALO_HOPS}, {&m->bwd2, false, "bwd2", 2 * HALO_HOPS}};
    for (auto& pl : plans) {
        if (!pl.P->n_chunks) { printf("  %-4s not available\n", pl.name); continue; }
        const Stats s = check_plan(m, *pl.P, pl.fwd, pl.name, pl.hops);
        printf("  %-4s tiles=%llu core/tile=%.1f occupied=%.3f plain_tiles=%.3f extras/tile=%.2f\n", pl.name, (unsigned long long)s.tiles, (double)s.core / s.tiles,
               (double)s.positions / (s.tiles * DENSE_LMAX), (double)s.plain / s.tiles, (double)s.extras / s.tiles);
    }
    CHECK(model_ensure_roi(m) == DBGPHMM_OK, "roi: %s", dbgphmm_last_error());
    for (int d = 0; d < 2; d++) {
        const uint32_t T = d ? m->bwd.n_chunks : m->fwd.n_chunks;
        const uint32_t* off = d ? m->d_roi_off_b : m->d_roi_off; const uint32_t* tl = d ? m->d_roi_tile_b : m->d_roi_tile; const uint32_t* tof = d ? m->d_tile_of_b : m->d_tile_of;
        for (uint32_t t = 0; t < T; t++) { CHECK(off[t] < off[t + 1] && tl[off[t]] == t, "roi list of tile %u", t); for (uint32_t a = off[t]; a < off[t + 1]; a++) CHECK(tl[a] < T, "roi tile id"); }
        for (uint32_t q = 0; q < N; q++) CHECK(tof[q] < T, "tile_of");
    }
    dbgphmm_model_destroy(m);
    return 0;
}

static int run_cache() {
    stub_device_bytes = 64ull << 20;
    cudaStream_t A, B; cudaStreamCreateWithFlags(&A, 0); cudaStreamCreateWithFlags(&B, 0);
    cache_set_stream(A);
    void* a = cache_alloc(3 << 20); void* b = cache_alloc(1000);
    CHECK(a && b && a != b && cache_unused_bytes() == 0, "fresh blocks");
    cache_free(a);
    CHECK(cache_unused_bytes() == (3u << 20), "unused bytes after a free");
    size_t waits = stub_stream_waits;
    void* a2 = cache_alloc((3 << 20) - 4096);            // same stream, fits within 1.5 x: the same block, no wait
    CHECK(a2 == a && stub_stream_waits == waits, "reuse on the freeing stream");
    cache_free(a2);
    cache_set_stream(B);
    void* a3 = cache_alloc(3 << 20);                     // another stream takes it behind the event of the free
    CHECK(a3 == a && stub_stream_waits == waits + 1, "reuse on another stream waits for the free");
    void* c = cache_alloc(1 << 20);                      // nothing unused fits: a new block
    CHECK(c && c != a && c != b, "no block to reuse");
    cache_free(a3); cache_free(c);
    void* d = cache_alloc(1 << 20);                      // best fit: the 1 MiB block, not the 3 MiB one (> 1.5 x + 1 MiB)
    CHECK(d == c, "best fit");
    cache_free(d);
    // a large request may take a block up to 4 x its size instead of allocating beside it
    void* big = cache_alloc(40ull << 20);
    CHECK(big, "40 MiB block");
    cache_free(big);
    const size_t allocs = stub_allocs;
    void* mid = cache_alloc(16ull << 20);                // < 64 MiB request: not eligible for oversize reuse
    CHECK(mid && mid != big && stub_allocs == allocs + 1, "requests under 64 MiB do not take oversize blocks");
    cache_free(mid);
    // out of memory: unused blocks are given back and the allocation is retried
    void* huge = cache_alloc(60ull << 20);
    CHECK(huge && cache_unused_bytes() == 0, "trim and retry");
    CHECK(cache_alloc(30ull << 20) == nullptr, "over capacity");
    cache_free(huge);
    // a free without a current stream: the next taker synchronises the device
    cache_set_stream(nullptr);
    void* e = cache_alloc(60ull << 20);
    CHECK(e == huge, "reuse without a stream");
    cache_free(e);
    cache_set_stream(A);
    const size_t syncs = stub_device_syncs;
    void* e2 = cache_alloc(60ull << 20);
    CHECK(e2 == huge && stub_device_syncs == syncs + 1, "block freed without a stream: device synchronisation before reuse");
    cache_free(e2); cache_free(b);
    // budget: 88 % of free + unused cache, unless fixed
    dbgphmm_model fake; fake.mem_budget_fixed = false;
    const uint64_t bud = model_budget(&fake);
    CHECK(bud == (uint64_t)((double)(stub_device_bytes - stub_bytes_in_use() + cache_unused_bytes()) * 0.88), "budget");
    fake.mem_budget_fixed = true; fake.mem_budget = 12345;
    CHECK(model_budget(&fake) == 12345, "fixed budget");
    cache_trim();
    CHECK(cache_unused_bytes() == 0 && stub_bytes_in_use() == 0 && stub_allocs == stub_frees, "everything returned");
    cache_set_stream(nullptr);
    cudaStreamDestroy(A); cudaStreamDestroy(B);
    printf("cache ok\n");
    return 0;
}

int main(int argc, char** argv) {
    if (const char* e = getenv("STUB_DEVICE_BYTES")) stub_device_bytes = strtoull(e, nullptr, 10);   // (default 1 GiB)
    if (argc >= 2 && !strcmp(argv[1], "cache")) return run_cache();
    if (argc >= 3 && !strcmp(argv[1], "plan")) { for (int i = 2; i < argc; i++) if (run_plan(argv[i])) return 1; return 0; }
    fprintf(stderr, "usage: host_logic plan <graph.bin>... | cache\n");
    return 2;
}
