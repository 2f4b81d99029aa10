// sparse.cu — sparse DP rows: per-row active-node selection + f_step / b_step over <= MAX_ACTIVE nodes.
//
// One CTA owns one (read, candidate X) job and walks its rows sequentially; the previous and the current row
// (SparseVec-like entry tables with a hash index) live in shared memory.  Reproduces, index-for-index:
//   PHMMTable::to_nodevec / top_nodes / top_nodes_by_score_ratio / filled_nodes   table.rs:117-149,199-211
//   to_childs, to_childs_and_us, to_parents_and_us (first-seen unique, cap 400)    active_nodes.rs:15-56
//   f_step with adaptive Del sets                                                  forward.rs:276-306,423-466
//   b_step with adaptive Del sets                                                  backward.rs:216-261,299-343
// Entry order of a row = insertion order of the reference's SparseVec: the step's `nodes` first (they hold m, i),
// then nodes that only ever received a Del value, in the order the Del rounds first touched them.
#include <algorithm>
#include <cstdlib>
#include "sparse.h"

#define SP_TENT 0x80000000u
#define SP_ABSENT 0xffffffffu

struct SGraph {
    uint32_t N, E;
    const uint8_t* emission;
    const double *init, *trans;
    const uint32_t *par_off, *par_node, *par_eid, *chi_off, *chi_node, *chi_eid;
    const uint32_t* pos_of;
    const uint4 *par_rec, *chi_rec;   // {CSR offset, degree, first neighbour, its edge id} per node (model.h)
};

// Shared-memory view of one job.  Every array sits at a fixed multiple of `cap` bytes from `base`, so the view is five
// registers and an array address is one multiply-add (a struct of ~35 pointers ends up in local memory as soon as it is
// passed by reference, and every access becomes a long-scoreboard local load).  cap must be a multiple of 16.
struct SS {
    unsigned char* base;
    uint32_t cap, hcap, hmask;
    int hshift;
    uint32_t htog;          // which of the two hash tables belongs to the current row
    uint32_t tog, ecall;    // block_prefix / sp_expand call parities
#define SS_ARR(type, name, off_in_caps) __device__ __forceinline__ type* name() const { return (type*)(base + (size_t)(off_in_caps) * cap); }
    // previous row (packed cells)
    SS_ARR(double, p_m, 0) SS_ARR(double, p_i, 8) SS_ARR(double, p_d, 16)
    // current row under construction
    SS_ARR(double, c_m, 24) SS_ARR(double, c_i, 32) SS_ARR(double, c_dv, 40)
    // Del values of the last two rounds ; stamp = round + 1 (entries are rebuilt every row)
    __device__ __forceinline__ double* dval(int k) const { return (double*)(base + (size_t)(48 + 8 * k) * cap); }
    SS_ARR(uint32_t, p_id, 64) SS_ARR(int, p_ex, 68)
    SS_ARR(uint32_t, c_id, 72) SS_ARR(int, c_mie, 76) SS_ARR(int, c_de, 80)
    __device__ __forceinline__ int* dexp(int k) const { return (int*)(base + (size_t)(84 + 4 * k) * cap); }
    SS_ARR(uint32_t, firstpos, 92)
    // ping-pong node lists (ids + slots) and the step's `nodes` (m/i entries)
    __device__ __forceinline__ uint32_t* la_id(int k) const { return (uint32_t*)(base + (size_t)(96 + 4 * k) * cap); }
    SS_ARR(uint32_t, act_id, 104)
    // the step's top list is dead once `nodes` has been expanded from it, before Del round 1 first writes la_id(1): same storage
    __device__ __forceinline__ uint32_t* top_id() const { return la_id(1); }
    __device__ __forceinline__ uint16_t* la_slot(int k) const { return (uint16_t*)(base + (size_t)(108 + 2 * k) * cap); }
    SS_ARR(uint16_t, act_slot, 112)
    SS_ARR(uint16_t, dlist, 114)         // slots in d insertion order
    __device__ __forceinline__ uint8_t* dstamp(int k) const { return (uint8_t*)(base + (size_t)(116 + k) * cap); }
    SS_ARR(uint8_t, d_seen, 118)
    SS_ARR(uint32_t, scan, 119)          // [cap + 4] scratch
    // ranking keys alias the Del round buffers, which are dead between rows
    __device__ __forceinline__ int* k_T() const { return dexp(0); }
    __device__ __forceinline__ unsigned long long* k_mant() const { return (unsigned long long*)dval(0); }
    // after the cap-proportional part (123 cap + 16 bytes): block_prefix scratch [2][8] + flags, then the two hash tables
    __device__ __forceinline__ uint32_t* wt() const { return (uint32_t*)(base + (size_t)123 * cap + 16); }
    __device__ __forceinline__ uint32_t* htab(uint32_t which) const { return (uint32_t*)(base + (size_t)123 * cap + 16 + 96) + (size_t)which * 2 * hcap; }
    __device__ __forceinline__ uint32_t* ch_key() const { return htab(htog); }
    __device__ __forceinline__ uint32_t* ch_val() const { return htab(htog) + hcap; }
    __device__ __forceinline__ uint32_t* ph_key() const { return htab(htog ^ 1u); }
    __device__ __forceinline__ uint32_t* ph_val() const { return htab(htog ^ 1u) + hcap; }
#undef SS_ARR
};
#define SS_BYTES(cap, hcap) ((size_t)123 * (cap) + 16 + 96 + (size_t)16 * (hcap))

__device__ __forceinline__ uint32_t sp_hash(uint32_t id, int shift) { return (id * 2654435761u) >> shift; }
__device__ __forceinline__ int sp_find(const uint32_t* key, const uint32_t* val, uint32_t hmask, int hshift, uint32_t id) {
    uint32_t h = sp_hash(id, hshift);
    for (uint32_t tries = 0; tries <= hmask; tries++) {   // (bounded: a completely full table has no empty cell to stop at)
        uint32_t k = key[h];
        if (k == id + 1) { uint32_t v = val[h]; return v < SP_TENT ? (int)v : -1; }
        if (k == 0) return -1;
        h = (h + 1) & hmask;
    }
    return -1;
}
// returns the cell holding `id`, inserting the key (value untouched = SP_ABSENT) if new ; SP_ABSENT if the table is full (a row on
// a heavily branching graph can offer more distinct candidates than 2 x cap cells: the caller then reports the row as too large for
// this capacity instead of probing forever)
__device__ __forceinline__ uint32_t sp_cell(uint32_t* key, uint32_t hmask, int hshift, uint32_t id) {
    uint32_t h = sp_hash(id, hshift);
    for (uint32_t tries = 0; tries <= 2 * hmask + 1; tries++) {
        uint32_t k = key[h];
        if (k == id + 1) return h;
        if (k == 0) {
            uint32_t old = atomicCAS(&key[h], 0u, id + 1);
            if (old == 0u || old == id + 1) return h;
        } else h = (h + 1) & hmask;
    }
    return SP_ABSENT;
}

// Ordered exclusive prefix of one small value per thread over the block (thread order) ; *total = block sum.
// One barrier: warp shuffles + per-warp totals in shared memory (double-buffered, so back-to-back calls need no
// second barrier).  All threads must call.  Two 16-bit counters may be packed into v.  At most 8 warps.
__device__ __forceinline__ uint32_t block_prefix(SS& S, uint32_t v, uint32_t* total) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    uint32_t x = v;   // becomes the inclusive prefix within the warp
    if (__all_sync(0xffffffffu, (v & 0xfffefffeu) == 0u)) {   // at most one per 16-bit counter (chains: degree 1): two ballots, no shuffle chain
        const uint32_t b0 = __ballot_sync(0xffffffffu, v & 1u), b1 = __ballot_sync(0xffffffffu, v >> 16);
        const uint32_t le = 0xffffffffu >> (31 - lane);
        x = (uint32_t)__popc(b0 & le) | ((uint32_t)__popc(b1 & le) << 16);
    } else {
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { uint32_t y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
    }
    uint32_t* wt = S.wt() + 8 * (S.tog & 1);
    S.tog++;
    if (lane == 31) wt[w] = x;
    __syncthreads();
    if (blockDim.x <= 64) {   // one or two warps (the common launch shapes): no loop over warp totals
        const uint2 a = *(const uint2*)wt;   // (the total of an absent warp stays 0)
        *total = a.x + a.y;
        return (w ? a.x : 0u) + x - v;
    }
    const uint4 a = *(const uint4*)wt, b = *(const uint4*)(wt + 4);   // totals of absent warps stay 0
    const uint32_t t[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    uint32_t base = 0, tot = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) { base += k < w ? t[k] : 0u; tot += t[k]; }
    *total = tot;
    return base + x - v;
}

__device__ __forceinline__ XF block_xsum(XF a) {  // deterministic block reduction; result valid in every thread
    __shared__ XF wv[32];
    __shared__ XF res;
    for (int o = 16; o; o >>= 1) {
        XF b; b.v = __shfl_down_sync(0xffffffffu, a.v, o); b.e = __shfl_down_sync(0xffffffffu, a.e, o);
        a = xadd(a, b);
    }
    if ((threadIdx.x & 31) == 0) wv[threadIdx.x >> 5] = a;
    __syncthreads();
    if (threadIdx.x == 0) {
        XF t = xf_zero();
        for (int k = 0; k < (int)((blockDim.x + 31) >> 5); k++) t = xadd(t, wv[k]);
        res = t;
    }
    __syncthreads();
    XF r = res;
    __syncthreads();
    return r;
}

// Unique, first-occurrence-ordered expansion of `src` (node ids) over an adjacency CSR (active_nodes.rs:15-56).
// and_us: the sources themselves come first.  with_nbrs: append their neighbours (in CSR = newest-edge-first order).
// Nodes without a current-row entry get one appended (zeroed).  Output: out_id/out_slot, *n_out (<= max_out).
// Returns false if the entry table would overflow `cap` (or a node has more than 16 neighbours).
// The candidate list [sources] ++ neighbours of source 0, 1, ... is never materialised: the thread of source q owns
// candidate q and the candidates base(q) .. base(q) + deg(q), base = ordered block prefix of the degrees.
//   phase 1  every candidate finds / inserts its hash cell and min-reduces its position into it
//   phase 2  a candidate is kept iff it holds the first position of its id ; ordered compaction by a second prefix,
//            sources first, then neighbours
__device__ bool sp_expand(SS& S, const uint32_t* src, int n_src, const uint4* rec, const uint32_t* nbr, bool and_us, bool with_nbrs,
                          uint32_t* out_id, uint16_t* out_slot, int max_out, uint32_t* n_ent_io, int* n_out) {
    __shared__ uint32_t s_flags[2][2];   // [call parity][0: overflow, 1: new entries not emitted]
    const int tid = threadIdx.x, B = blockDim.x;
    const uint32_t n_ent0 = *n_ent_io;
    const bool cap_limited = (int)S.cap < max_out;  // the table cannot hold the reference's 400-entry list
    if (cap_limited) max_out = (int)S.cap;
    uint32_t* fl = s_flags[S.ecall & 1];
    S.ecall++;
    for (uint32_t e = tid; e < n_ent0; e += B) S.firstpos()[e] = SP_ABSENT;
    if (tid == 0) { fl[0] = 0; fl[1] = 0; }
    const uint32_t n_self = and_us ? (uint32_t)n_src : 0u;
    auto note = [&](uint32_t id, uint32_t p) -> uint32_t {   // phase 1 for one candidate ; returns its hash cell
        const uint32_t cell = sp_cell(S.ch_key(), S.hmask, S.hshift, id);
        if (cell == SP_ABSENT) { fl[0] = 1; return 0; }       // hash table full: the row does not fit this capacity
        const uint32_t v = S.ch_val()[cell];
        if (v < SP_TENT) atomicMin(&S.firstpos()[v], p);
        else atomicMin(&S.ch_val()[cell], SP_TENT | p);
        return cell;
    };
    if (!with_nbrs) __syncthreads();   // (the prefix below is the barrier otherwise) firstpos reset before the min-reductions
    uint32_t C = n_self;
    for (int q0 = 0; q0 < n_src; q0 += B) {
        const int q = q0 + tid;
        uint32_t id = 0, cnt = 0;
        uint4 r = make_uint4(0u, 0u, 0u, 0u);
        if (q < n_src) { id = src[q]; if (with_nbrs) { r = __ldg(rec + id); cnt = r.y; } }
        if (with_nbrs) {
            const bool too_many = cnt > 16;
            if (too_many) cnt = 0;
            uint32_t tot;
            const uint32_t base = C + block_prefix(S, cnt, &tot);
            C += tot;
            if (q < n_src) {
                if (too_many) fl[0] = 1;
                uint32_t cell0 = 0;
                for (uint32_t k = 0; k < cnt; k++) { const uint32_t c = note(k ? nbr[r.x + k] : r.z, base + k); if (k == 0) cell0 = c; }
                // position of the first candidate (< 8192: at most 400 + 16 x 400), degree, and the hash cell of the first neighbour
                // (phase 2 does not probe for it again)
                S.scan()[q] = base | (cnt << 13) | (cell0 << 18);
            }
        }
        if (and_us && q < n_src) note(id, (uint32_t)q);
    }
    __syncthreads();
    if (fl[0]) { __syncthreads(); return false; }   // (uniform: some candidate found no hash cell or a node has too many neighbours ; phase 2 would probe for keys that are not there)
    // kept / new flag of candidate p with id `id` : bit 0 kept, bit 16 new entry
    auto cell_of = [&](uint32_t id) -> uint32_t {
        uint32_t h = sp_hash(id, S.hshift);
        for (uint32_t tries = 0; tries <= S.hmask && S.ch_key()[h] != id + 1; tries++) h = (h + 1) & S.hmask;   // inserted in phase 1
        return h;
    };
    auto flag_at = [&](uint32_t cell, uint32_t p, uint32_t* v_out) -> uint32_t {
        const uint32_t v = S.ch_val()[cell];
        *v_out = v;
        if (v < SP_TENT) return (S.firstpos()[v] == p) ? 1u : 0u;
        return v == (SP_TENT | p) ? 0x10001u : 0u;
    };
    auto emit = [&](uint32_t id, uint32_t p, uint32_t f, uint32_t excl, uint32_t cell, uint32_t v) {
        const bool isnew = f >> 16;
        const uint32_t oi = excl & 0xffffu, ni = excl >> 16;
        const uint32_t slot = isnew ? n_ent0 + ni : v;
        const bool em = (int)oi < max_out;
        if (em) {
            if (slot >= S.cap) fl[0] = 1;
            else {
                if (isnew) {
                    S.c_id()[slot] = id; S.c_m()[slot] = 0.0; S.c_i()[slot] = 0.0; S.c_mie()[slot] = 0; S.c_dv()[slot] = 0.0; S.c_de()[slot] = XF_ZERO_E;
                    S.dstamp(0)[slot] = 0; S.dstamp(1)[slot] = 0; S.d_seen()[slot] = 0; S.firstpos()[slot] = p;
                }
                out_id[oi] = id; out_slot[oi] = (unsigned short)slot;
            }
        }
        if (isnew) {
            S.ch_val()[cell] = (em && slot < S.cap) ? slot : SP_ABSENT;
            if (!em) atomicAdd(&fl[1], 1u);
        }
    };
    uint32_t run = 0;   // kept count in the low half, new-entry count in the high half
    if (and_us) {
        for (int q0 = 0; q0 < n_src; q0 += B) {
            const int q = q0 + tid;
            uint32_t id = 0, f = 0, cell = 0, v = 0;
            if (q < n_src) { id = src[q]; cell = cell_of(id); f = flag_at(cell, (uint32_t)q, &v); }
            uint32_t tot;
            const uint32_t excl = run + block_prefix(S, f, &tot);   // (barrier: every tentative value of this chunk has been read)
            if (f) emit(id, (uint32_t)q, f, excl, cell, v);
            run += tot;
            if (with_nbrs || q0 + B < n_src) __syncthreads();   // later candidates must see the slots published by this chunk
        }
    }
    if (with_nbrs) {
        for (int q0 = 0; q0 < n_src; q0 += B) {
            const int q = q0 + tid;
            uint32_t base = 0, cnt = 0, cell0 = 0, mask = 0, mine = 0;
            uint4 r = make_uint4(0u, 0u, 0u, 0u);
            if (q < n_src) {
                const uint32_t sc = S.scan()[q];
                base = sc & 0x1fffu; cnt = (sc >> 13) & 31u; cell0 = sc >> 18;
                if (cnt) r = __ldg(rec + src[q]);
                for (uint32_t k = 0; k < cnt; k++) {
                    uint32_t v;
                    const uint32_t f = flag_at(k ? cell_of(nbr[r.x + k]) : cell0, base + k, &v);
                    if (f) { mask |= (f >> 16 ? 3u : 1u) << (2 * k); mine += f; }
                }
            }
            uint32_t tot;
            uint32_t excl = run + block_prefix(S, mine, &tot);
            for (uint32_t k = 0; k < cnt; k++) {
                const uint32_t b2 = (mask >> (2 * k)) & 3u;
                if (!b2) continue;
                const uint32_t id = k ? nbr[r.x + k] : r.z, f = b2 == 3u ? 0x10001u : 1u;
                const uint32_t cell = k ? cell_of(id) : cell0;
                const uint32_t v = S.ch_val()[cell];   // (the slot again ; the value is not published yet: only this thread does)
                emit(id, base + k, f, excl, cell, v);
                excl += f;
            }
            run += tot;
            if (q0 + B < n_src) __syncthreads();
        }
    }
    __syncthreads();
    const uint32_t kept_tot = run & 0xffffu, new_tot = run >> 16;
    *n_out = (int)(kept_tot < (uint32_t)max_out ? kept_tot : (uint32_t)max_out);
    *n_ent_io = n_ent0 + new_tot - fl[1];
    return !fl[0] && !(cap_limited && kept_tot > (uint32_t)max_out);
}

// merged value key of packed entry e of the previous row
__device__ __forceinline__ void sp_key(const SS& S, uint32_t e, int* T, unsigned long long* mant) {
    double v = S.p_m()[e] + S.p_i()[e] + S.p_d()[e];
    if (v == 0.0) { *T = XF_ZERO_E; *mant = 0; return; }
    long long b = __double_as_longlong(v);
    *T = S.p_ex()[e] + (int)((b >> 52) & 0x7ff) - 1023;
    *mant = (unsigned long long)b & 0xfffffffffffffull;
}

// Ranking keys.  (T, 52-bit mantissa) of a merged value ; zero value: T == XF_ZERO_E.  When every T lies within +-1000 binades of
// Tref (practically always) the pair packs into one 64-bit integer and a comparison is one instruction pair ; `bad` is raised
// otherwise and the ranking falls back to the two-word comparison (the mantissa is the low 52 bits either way).
__device__ __forceinline__ unsigned long long sp_pack(int T, unsigned long long mant, int Tref, uint32_t* bad) {
    if (T == XF_ZERO_E) return 0;
    const int rel = T - Tref + 1024;
    if (rel < 1 || rel > 2046) *bad = 1;
    return ((unsigned long long)(unsigned)rel << 52) | mant;
}
// Rank n keys (k_T, k_mant packed by sp_pack) in descending order, ties by position (UNPINNED in the reference):
// out_id[rank] = ids[e] for rank < KK and scan[e] = rank.  Caller: barrier before (keys written) ; ends with a barrier.
__device__ void sp_rank(SS& S, uint32_t n, uint32_t KK, const uint32_t* ids, uint32_t* out_id) {
    const int tid = threadIdx.x, B = blockDim.x;
    uint32_t* bad = S.wt() + 16;   // zero between rankings (reset below after use)
    if (!*bad) {
        // f precedes e iff kf > ke, or kf == ke and f < e: that is kf + (f < e) > ke (keys stay below 2^63) ; two keys per load
        // (a variant with thresholds ke - (f < e) and warp-uniform loop ranges measured 5 % slower: three short loops instead of one)
        const ulonglong2* k2 = (const ulonglong2*)S.k_mant();   // (16-byte aligned: cap is a multiple of 16)
        for (uint32_t e = tid; e < n; e += B) {
            const unsigned long long ke = S.k_mant()[e];
            uint32_t rank = 0;
#pragma unroll 2
            for (uint32_t f = 0; f + 1 < n; f += 2) {
                const ulonglong2 kf = k2[f >> 1];
                rank += (kf.x + (f < e ? 1ull : 0ull) > ke) ? 1u : 0u;
                rank += (kf.y + (f + 1 < e ? 1ull : 0ull) > ke) ? 1u : 0u;
            }
            if (n & 1u) rank += (S.k_mant()[n - 1] + (n - 1 < e ? 1ull : 0ull) > ke) ? 1u : 0u;
            if (rank < KK) out_id[rank] = ids[e];
            S.scan()[e] = rank;
        }
    } else {
        const unsigned long long M52 = 0xfffffffffffffull;
        for (uint32_t e = tid; e < n; e += B) {
            const int T = S.k_T()[e]; const unsigned long long mt = T == XF_ZERO_E ? 0ull : (S.k_mant()[e] & M52);
            uint32_t rank = 0;
            for (uint32_t f = 0; f < n; f++) {
                const int Tf = S.k_T()[f]; const unsigned long long mf = Tf == XF_ZERO_E ? 0ull : (S.k_mant()[f] & M52);
                rank += ((Tf > T) || (Tf == T && (mf > mt || (mf == mt && f < e)))) ? 1u : 0u;
            }
            if (rank < KK) out_id[rank] = ids[e];
            S.scan()[e] = rank;
        }
    }
    __syncthreads();
    if (tid == 0) *bad = 0;
}

// top-K ids of the previous (packed) row in descending merged value, ties by entry position (UNPINNED in the
// reference); by_ratio keeps the prefix with ln v0 - ln v < ratio of the top-400 list (table.rs:134-149).
__device__ int sp_top_of_prev(SS& S, uint32_t n_prev, uint32_t K, bool by_ratio, double ratio, uint32_t* out_id) {
    __shared__ double s_L0;
    __shared__ uint32_t s_cnt;
    const int tid = threadIdx.x, B = blockDim.x;
    const int Tref = S.p_ex()[0];
    for (uint32_t e = tid; e < n_prev; e += B) {
        int T; unsigned long long mant;
        sp_key(S, e, &T, &mant);
        S.k_T()[e] = T; S.k_mant()[e] = sp_pack(T, mant, Tref, S.wt() + 16);
    }
    if (tid == 0) { s_cnt = 0; s_L0 = -INFINITY; }
    __syncthreads();
    const uint32_t KK = K < n_prev ? K : n_prev;
    sp_rank(S, n_prev, KK, S.p_id(), out_id);
    if (!by_ratio) return (int)KK;
    for (uint32_t e = tid; e < n_prev; e += B)
        if (S.scan()[e] == 0) s_L0 = xlog(xf(S.p_m()[e] + S.p_i()[e] + S.p_d()[e], S.p_ex()[e]));
    __syncthreads();
    for (uint32_t e = tid; e < n_prev; e += B) {
        if (S.scan()[e] < KK) {
            double L = xlog(xf(S.p_m()[e] + S.p_i()[e] + S.p_d()[e], S.p_ex()[e]));
            if (s_L0 - L < ratio) atomicAdd(&s_cnt, 1u);
        }
    }
    __syncthreads();
    int r = (int)s_cnt;
    __syncthreads();
    return r;
}

struct PrevAcc {  // accessor of the row before the current one
    int kind;     // 0 sparse (shared memory), 1 dense slab, 2 B-init, 3 F-init, 4 gathered cells of a dense row
    const double *gm, *gi, *gd; const int* ge;
    const uint32_t* gid; uint32_t gn;   // kind 4: ids of the gathered cells (unsorted, duplicates allowed) and their number
    double p_end;
};
__device__ __forceinline__ void prev_get(const SS& S, const PrevAcc& P, uint32_t id, double* m, double* i, double* d, int* ex) {
    if (P.kind == 0) {
        int sl = sp_find(S.ph_key(), S.ph_val(), S.hmask, S.hshift, id);
        if (sl < 0) { *m = *i = *d = 0.0; *ex = 0; } else { *m = S.p_m()[sl]; *i = S.p_i()[sl]; *d = S.p_d()[sl]; *ex = S.p_ex()[sl]; }
    } else if (P.kind == 1) { *m = P.gm[id]; *i = P.gi[id]; *d = P.gd[id]; *ex = P.ge[id]; }
    else if (P.kind == 2) { *m = *i = *d = P.p_end; *ex = 0; }
    else if (P.kind == 4) {   // linear search: this is one row per job, and the list holds every cell that row can ask for
        *m = *i = *d = 0.0; *ex = 0;
        for (uint32_t e = 0; e < P.gn; e++)
            if (P.gid[e] == id) { *m = P.gm[e]; *i = P.gi[e]; *d = P.gd[e]; *ex = P.ge[e]; break; }
    }
    else { *m = *i = *d = 0.0; *ex = 0; }
}

extern __shared__ __align__(16) unsigned char sp_smem[];

// role 0, no rescue queue: CTA b runs job b ; a job whose row outgrows the launch's entry capacity reports SJ_NEED_BIG.
// role 0 with a rescue queue: persistent CTAs take jobs off a counter ; a job whose row outgrows the capacity hands its previous
// row over to the queue and the CTA moves on.  role 1: persistent CTAs of a second launch with the next capacity, running
// beside the primary one, take such jobs from the queue and carry them on from the row that did not fit.  (Re-running a failed
// job in a later pass would cost a whole extra job latency: the kernel is latency-bound, one straggler costs as much as a wave.
// Launch order primary, rescue: neither waits for the other to start, so the pair also completes when kernels are serialised.)
// DIR: 0 every job of the launch runs forward, 1 backward (the launches of a phase are uniform ; a specialised kernel is half the
// code of the generic one, and the kernel is sensitive to instruction-cache misses: 16 warps per SM wander through > 100 KB of it).
template <int DIR>
__global__ void k_sparse(SGraph G, LinParams lp, const SJob* __restrict__ jobs, SparseIO io, uint32_t cap, uint32_t hcap, int role) {
    const int tid = threadIdx.x, B = blockDim.x;
    // ---- shared-memory view
    SS S;
    S.base = sp_smem; S.cap = cap; S.hcap = hcap; S.hmask = hcap - 1; S.hshift = 32 - (31 - __clz(hcap));
    S.htog = 0; S.tog = 0; S.ecall = 0;
    if (threadIdx.x < 24) S.wt()[threadIdx.x] = 0;

    __shared__ XF s_mb, s_ib;          // begin scalars of the previous row (forward) / ib of the next row (backward)
    __shared__ uint64_t s_page_off;   // start of a freshly taken arena page (thread 0 -> block)
    __shared__ int s_fail;
    __shared__ unsigned long long s_cells;
    __shared__ unsigned long long s_info;   // where a job stopped: step << 32 | site << 16 | entries
    __shared__ uint32_t s_job;
  for (;;) {
    uint32_t job_idx = blockIdx.x, s_begin = 0;
    if (role == 0 && io.rq_ctl) {   // persistent primary CTAs: next job of the launch
        if (tid == 0) s_job = atomicAdd(io.rq_ctl + 3, 1u);
        __syncthreads();
        job_idx = s_job;
        __syncthreads();
        if (job_idx >= io.rq_n_jobs) {
            if (tid == 0) { __threadfence(); atomicAdd(io.rq_ctl + 2, 1u); }
            return;
        }
    }
    if (role == 1) {   // take the next handed-over job, or leave once every primary CTA is gone and the queue is empty
        if (tid == 0) {
            volatile uint32_t* ctl = io.rq_ctl;
            uint32_t got = SP_ABSENT;
            // No fence while waiting: a gpu-scope fence invalidates the whole L1 of the SM (CCTL.IVALL), and this CTA shares its SM
            // with the primary CTAs, whose graph reads live there.  The counters are read with volatile (L1-bypassing) loads ; the
            // one decision that needs ordering -- leaving -- fences once: a push always precedes its CTA's done mark, so the counters
            // read after "every primary CTA is gone" + fence are final.
            for (;;) {
                const uint32_t pushed = ctl[0], popped = ctl[1];
                if (popped < pushed) { if (atomicCAS(io.rq_ctl + 1, popped, popped + 1) == popped) { got = popped; break; } }
                else if (ctl[2] == io.rq_n_primary) {
                    __threadfence();
                    if (ctl[1] >= ctl[0]) break;
                } else __nanosleep(2000);
            }
            if (got != SP_ABSENT) {
                volatile uint32_t* it = io.rq_items + got;
                uint32_t j;
                while ((j = *it) == SP_ABSENT) __nanosleep(200);
                __threadfence();
                got = j;
            }
            s_job = got;
        }
        __syncthreads();
        job_idx = s_job;
        __syncthreads();
        if (job_idx == SP_ABSENT) return;
    }
    const SJob jb = jobs[job_idx];
    const bool skip = (jb.active_idx >= 0 && !io.active[jb.active_idx]) || jb.n_rows == 0;
    const double* init = G.init + (size_t)jb.x * G.N;
    const double* trans = G.trans + (size_t)jb.x * G.E;
    constexpr bool fwd = DIR == 0;
    const bool adaptive = (jb.mode == SP_TOPN || jb.mode == SP_RATIO);
    uint32_t n_prev = 0;   // packed entries of the previous row (sparse prev only)
    uint64_t page_off = 0, page_left = 0;   // arena page this job is filling (block-uniform registers)
    XF last_scalar = xf_zero();
    if (tid == 0) {
        s_page_off = 0; s_fail = SJ_OK; s_cells = 0; s_info = 0;
        if (fwd) {
            if (jb.row_begin == 0) { s_mb = xf(1.0, 0); s_ib = xf_zero(); }
            else { s_mb = io.desc[jb.desc0 + jb.row_begin - 1].mb; s_ib = io.desc[jb.desc0 + jb.row_begin - 1].ib; }
        } else {
            s_mb = xf_zero();
            s_ib = (jb.row_begin == (int)jb.len - 1) ? xf_zero() : io.desc[jb.desc0 + jb.row_begin + 1].ib;
        }
    }
    if (role == 1) {   // resume: the row before step s_begin as the primary CTA left it
        SHandoff h;
        {   // (written by another SM: read through L2)
            const uint4* hp = (const uint4*)(io.rq_hand + job_idx); uint4* hq = (uint4*)&h;
#pragma unroll
            for (int k = 0; k < (int)(sizeof(SHandoff) / 16); k++) hq[k] = __ldcg(hp + k);
        }
        s_begin = h.step; n_prev = h.n_prev; last_scalar = h.last_scalar;
        __syncthreads();
        if (tid == 0 && s_begin > 0) { s_mb = h.mb; s_ib = h.ib; s_cells = h.cells; }
        const char* hr = io.rq_rows + (size_t)job_idx * 32 * io.rq_cap;
        const double* hm = (const double*)hr; const double* hi = hm + io.rq_cap; const double* hd = hi + io.rq_cap;
        const uint32_t* hid = (const uint32_t*)(hd + io.rq_cap); const int* hex = (const int*)(hid + io.rq_cap);
        for (uint32_t h2 = tid; h2 < hcap; h2 += B) { S.ph_key()[h2] = 0; S.ph_val()[h2] = SP_ABSENT; }
        for (uint32_t e = tid; e < n_prev; e += B) { S.p_m()[e] = __ldcg(hm + e); S.p_i()[e] = __ldcg(hi + e); S.p_d()[e] = __ldcg(hd + e); S.p_id()[e] = __ldcg(hid + e); S.p_ex()[e] = __ldcg(hex + e); }
        __syncthreads();
        for (uint32_t e = tid; e < n_prev; e += B) S.ph_val()[sp_cell(S.ph_key(), S.hmask, S.hshift, S.p_id()[e])] = e;
    }
    __syncthreads();

    for (uint32_t s = s_begin; s < jb.n_rows && !skip; s++) {
        const int row = fwd ? jb.row_begin + (int)s : jb.row_begin - (int)s;
        const uint8_t x = io.bases[jb.base_off + row];
        PrevAcc PA;
        PA.p_end = lp.p_end; PA.gm = PA.gi = PA.gd = nullptr; PA.ge = nullptr; PA.gid = nullptr; PA.gn = 0;
        if (s > 0) PA.kind = 0;
        else if (jb.prev0_kind == SPREV_GATHER) {
            PA.kind = 4;
            const char* gl = io.gather + (size_t)jb.top0 * 32 * io.gather_cap;
            PA.gm = (const double*)gl; PA.gi = PA.gm + io.gather_cap; PA.gd = PA.gi + io.gather_cap;
            PA.gid = (const uint32_t*)(PA.gd + io.gather_cap); PA.ge = (const int*)(PA.gid + io.gather_cap);
            PA.gn = io.gather_cnt[jb.top0];
        }
        else if (jb.prev0_kind == SPREV_DENSE) {
            PA.kind = 1;
            const char* sl = io.pool + jb.prev0_slab * io.slab_bytes;
            PA.gm = (const double*)sl; PA.gi = PA.gm + io.Np; PA.gd = PA.gi + io.Np; PA.ge = (const int*)(PA.gd + io.Np);
        } else PA.kind = jb.prev0_kind == SPREV_B_INIT ? 2 : 3;

        // ---------------- 1. node list of this step
        int n_top = 0;
        if (jb.mode == SP_TOPN || jb.mode == SP_RATIO) {
            if (s == 0) {
                n_top = (int)io.top_cnt[jb.top0];
                if (n_top <= (int)cap) for (int t = tid; t < n_top; t += B) S.top_id()[t] = io.top_ids[(size_t)jb.top0 * MAX_ACTIVE + t];
            } else {
                n_top = sp_top_of_prev(S, n_prev, jb.mode == SP_RATIO ? MAX_ACTIVE : lp.n_active_nodes, jb.mode == SP_RATIO,
                                       lp.active_node_max_ratio, S.top_id());
            }
        } else if (jb.mode == SP_MAPPING) {
            uint64_t a = io.map_row_off[jb.map_row0 + row], b = io.map_row_off[jb.map_row0 + row + 1];
            n_top = (int)(b - a);
            if (b - a > MAX_ACTIVE) { if (tid == 0) s_fail = SJ_CAPACITY; n_top = 0; }   // (the reference's 400-entry SparseVec panics)
            if (n_top <= (int)cap) for (int t = tid; t < n_top; t += B) S.top_id()[t] = io.map_nodes[a + t];
        } else {  // SP_BYFWD: filled_nodes() of forward row (row-1): top-|m entries| of its merged vector
            const RowDesc fr = io.fdesc[jb.fdesc0 + row - 1];
            const char* pay = io.farena + fr.off;
            const double* fm = (const double*)pay; const double* fi = fm + fr.n_ent; const double* fd = fi + fr.n_ent;
            const uint32_t* fid = (const uint32_t*)(fd + fr.n_ent); const int* fex = (const int*)(fid + fr.n_ent);
            if (fr.n_ent > cap) { if (tid == 0) s_fail = SJ_NEED_BIG; }
            else {
                // stage into the *current* arrays (not yet in use) and rank there
                const int Tref = fr.n_ent ? fex[0] : 0;
                for (uint32_t e = tid; e < fr.n_ent; e += B) {
                    const double v = fm[e] + fi[e] + fd[e];
                    int T = XF_ZERO_E; unsigned long long mant = 0;
                    if (v != 0.0) { const long long bb = __double_as_longlong(v); T = fex[e] + (int)((bb >> 52) & 0x7ff) - 1023; mant = (unsigned long long)bb & 0xfffffffffffffull; }
                    S.k_T()[e] = T; S.k_mant()[e] = sp_pack(T, mant, Tref, S.wt() + 16);
                    S.c_id()[e] = fid[e];
                }
                __syncthreads();
                sp_rank(S, fr.n_ent, fr.n_mi, S.c_id(), S.top_id());
                n_top = (int)fr.n_mi;
            }
        }
        if (n_top > (int)cap && tid == 0) s_fail = SJ_NEED_BIG;   // list scratch is sized by cap
        __syncthreads();
        if (s_fail) { if (tid == 0) s_info = ((unsigned long long)s << 32) | 1u << 16 | (unsigned)n_top; break; }

        // ---------------- 2. reset the current row
        uint32_t n_ent = 0;
        {   // (16-byte stores: the tables are 16-byte aligned and hcap is a power of two >= 64)
            uint4* hk = (uint4*)S.ch_key(); uint4* hv = (uint4*)S.ch_val();
            const uint4 z = make_uint4(0u, 0u, 0u, 0u), ab = make_uint4(SP_ABSENT, SP_ABSENT, SP_ABSENT, SP_ABSENT);
            for (uint32_t h = tid; h < hcap / 4; h += B) { hk[h] = z; hv[h] = ab; }
        }
        __syncthreads();

        // ---------------- 3. the step's `nodes` (they hold m, i)
        int n_act = 0;
        bool ok = true;
        if (fwd) {
            // forward sparse: nodes = to_childs_and_us(top) (forward.rs:148) ; mapping: nodes = mapping.nodes(i)
            ok = sp_expand(S, S.top_id(), n_top, G.chi_rec, G.chi_node, true, adaptive, S.act_id(), S.act_slot(), MAX_ACTIVE, &n_ent, &n_act);
        } else {
            // backward sparse: M/I over to_parents_and_us(nodes) (backward.rs:243-259) ; non-adaptive: nodes themselves
            ok = sp_expand(S, S.top_id(), n_top, G.par_rec, G.par_node, true, adaptive, S.act_id(), S.act_slot(), MAX_ACTIVE, &n_ent, &n_act);
        }
        if (!ok) { if (tid == 0) { s_fail = SJ_NEED_BIG; s_info = ((unsigned long long)s << 32) | 2u << 16 | n_ent; } __syncthreads(); break; }
        const uint32_t n_mi = (uint32_t)n_act;
        const uint32_t stamp0 = 1;   // Del stamps: round t writes t + 1
        uint32_t n_d = 0;

        if (fwd) {
            // ---------------- 4f. fm, fi over nodes (forward.rs:337-388)
            const XF fb0 = xadd(xmul(s_mb, lp.p_MM), xmul(s_ib, lp.p_IM));
            const XF ib_cur = xmul(xadd(xmul(s_mb, lp.p_MI), xmul(s_ib, lp.p_II)), lp.p_random);
            for (uint32_t a = tid; a < n_mi; a += B) {
                uint32_t id = S.act_id()[a], sl = S.act_slot()[a];
                XF acc = xf_zero();
                const uint4 r = __ldg(G.par_rec + id);
                for (uint32_t k = 0; k < r.y; k++) {
                    const uint32_t pn = k ? G.par_node[r.x + k] : r.z, eid = k ? G.par_eid[r.x + k] : r.w;
                    const double tr = trans[eid];
                    double pm, pi, pd; int pe;
                    prev_get(S, PA, pn, &pm, &pi, &pd, &pe);
                    acc = xadd(acc, xf(tr * (lp.p_MM * pm + lp.p_IM * pi + lp.p_DM * pd), pe));
                }
                acc = xadd(acc, xmul(fb0, init[id]));
                XF m = xmul(acc, G.emission[id] == x ? lp.p_match : lp.p_mismatch);
                double pm, pi, pd; int pe;
                prev_get(S, PA, id, &pm, &pi, &pd, &pe);
                XF i = xf(lp.p_random * (lp.p_MI * pm + lp.p_II * pi + lp.p_DI * pd), pe);
                int Em = xexp(m), Ei = xexp(i), Ec = Em > Ei ? Em : Ei;
                if (Ec == XF_ZERO_E) { S.c_m()[sl] = 0.0; S.c_i()[sl] = 0.0; S.c_mie()[sl] = 0; }
                else { S.c_m()[sl] = m.v == 0.0 ? 0.0 : m.v * pow2i(m.e - Ec); S.c_i()[sl] = i.v == 0.0 ? 0.0 : i.v * pow2i(i.e - Ec); S.c_mie()[sl] = Ec; }
            }
            __syncthreads();
            // ---------------- 5f. fd: fd0 + 4 x fdt (forward.rs:423-466)
            const uint32_t* src_id = S.act_id(); int n_src = n_act;
            for (int t = 0; t < N_DEL_ROUNDS; t++) {
                const uint32_t* l_id = S.act_id(); const uint16_t* l_slot = S.act_slot();   // non-adaptive: every round runs over `nodes`
                int n_l = n_act;
                if (adaptive) {
                    ok = sp_expand(S, src_id, n_src, G.chi_rec, G.chi_node, false, true, S.la_id(t & 1), S.la_slot(t & 1), MAX_ACTIVE, &n_ent, &n_l);
                    if (!ok) break;
                    l_id = S.la_id(t & 1); l_slot = S.la_slot(t & 1);
                }
                double* dv = S.dval(t & 1); int* de = S.dexp(t & 1); uint8_t* st = S.dstamp(t & 1);
                const double* dvp = S.dval((t & 1) ^ 1); const int* dep = S.dexp((t & 1) ^ 1); const uint8_t* stp = S.dstamp((t & 1) ^ 1);
                for (int a0 = 0; a0 < n_l; a0 += B) {
                    const int a = a0 + tid;
                    uint32_t sl = 0, fresh = 0;
                    if (a < n_l) {
                        const uint32_t id = l_id[a];
                        sl = l_slot[a];
                        XF acc = xf_zero();
                        const uint4 r = __ldg(G.par_rec + id);
                        for (uint32_t k = 0; k < r.y; k++) {
                            const uint32_t pn = k ? G.par_node[r.x + k] : r.z, eid = k ? G.par_eid[r.x + k] : r.w;
                            const double tr = trans[eid];   // (issued before the hash probe: its latency overlaps it)
                            int ps = sp_find(S.ch_key(), S.ch_val(), S.hmask, S.hshift, pn);
                            if (ps < 0) continue;
                            if (t == 0) { if ((uint32_t)ps < n_mi) acc = xadd(acc, xf(tr * (lp.p_MD * S.c_m()[ps] + lp.p_ID * S.c_i()[ps]), S.c_mie()[ps])); }
                            else if (stp[ps] == stamp0 + t - 1) acc = xadd(acc, xf(tr * lp.p_DD * dvp[ps], dep[ps]));
                        }
                        if (t == 0) acc = xadd(acc, xmul(ib_cur, lp.p_ID * init[id]));
                        dv[sl] = acc.v; de[sl] = acc.e; st[sl] = (uint8_t)(stamp0 + t);
                        XF tot = xadd(xf(S.c_dv()[sl], S.c_de()[sl]), acc);
                        S.c_dv()[sl] = tot.v; S.c_de()[sl] = tot.e;
                        fresh = S.d_seen()[sl] ? 0u : 1u;
                    }
                    // d insertion order: first time a slot receives a Del value (non-adaptive: round 0, in list order)
                    if (adaptive) {
                        uint32_t n_new;
                        const uint32_t o = n_d + block_prefix(S, fresh, &n_new);
                        if (fresh) { if (o < cap) S.dlist()[o] = (uint16_t)sl; S.d_seen()[sl] = 1; }
                        n_d += n_new;
                    } else if (t == 0 && a < n_l) S.dlist()[a] = (uint16_t)sl;
                }
                if (!adaptive && t == 0) n_d = (uint32_t)n_l;
                __syncthreads();
                src_id = l_id; n_src = n_l;
            }
            if (!ok) { if (tid == 0) { s_fail = SJ_NEED_BIG; s_info = ((unsigned long long)s << 32) | 3u << 16 | n_ent; } __syncthreads(); break; }
            // ---------------- 6f. fe over nodes (forward.rs:554-558), begin scalars
            XF part = xf_zero();
            for (uint32_t a = tid; a < n_mi; a += B) {
                uint32_t sl = S.act_slot()[a];
                part = xadd(part, xadd(xf(S.c_m()[sl] + S.c_i()[sl], S.c_mie()[sl]), xf(S.c_dv()[sl], S.c_de()[sl])));
            }
            XF esum = block_xsum(part);
            last_scalar = xnorm(xmul(esum, lp.p_end));
            if (tid == 0) { s_mb = xf_zero(); s_ib = xnorm(ib_cur); }
        } else {
            // ---------------- 4b. bd: bd0 + 4 x bdt over iterated to_parents_and_us (backward.rs:299-343)
            const uint32_t* src_id = S.act_id(); int n_src = n_act;
            for (int t = 0; t < N_DEL_ROUNDS; t++) {
                const uint32_t* l_id; const uint16_t* l_slot; int n_l;
                if (t == 0 || !adaptive) { l_id = S.act_id(); l_slot = S.act_slot(); n_l = n_act; }  // A0 == to_parents_and_us(nodes)
                else {
                    int nn = 0;
                    ok = sp_expand(S, src_id, n_src, G.par_rec, G.par_node, true, true, S.la_id(t & 1), S.la_slot(t & 1), MAX_ACTIVE, &n_ent, &nn);
                    if (!ok) break;
                    l_id = S.la_id(t & 1); l_slot = S.la_slot(t & 1); n_l = nn;
                }
                double* dv = S.dval(t & 1); int* de = S.dexp(t & 1); uint8_t* st = S.dstamp(t & 1);
                const double* dvp = S.dval((t & 1) ^ 1); const int* dep = S.dexp((t & 1) ^ 1); const uint8_t* stp = S.dstamp((t & 1) ^ 1);
                for (int a0 = 0; a0 < n_l; a0 += B) {
                    const int a = a0 + tid;
                    uint32_t sl = 0, fresh = 0;
                    if (a < n_l) {
                        const uint32_t id = l_id[a];
                        sl = l_slot[a];
                        XF acc = xf_zero();
                        const uint4 r = __ldg(G.chi_rec + id);
                        for (uint32_t k = 0; k < r.y; k++) {
                            const uint32_t ch = k ? G.chi_node[r.x + k] : r.z, eid = k ? G.chi_eid[r.x + k] : r.w;
                            const double tr = trans[eid];
                            if (t == 0) {
                                double pm, pi, pd; int pe;
                                prev_get(S, PA, ch, &pm, &pi, &pd, &pe);
                                acc = xadd(acc, xf(tr * lp.p_DM * (G.emission[ch] == x ? lp.p_match : lp.p_mismatch) * pm, pe));
                            } else {
                                int cs = sp_find(S.ch_key(), S.ch_val(), S.hmask, S.hshift, ch);
                                if (cs >= 0 && stp[cs] == stamp0 + t - 1) acc = xadd(acc, xf(tr * lp.p_DD * dvp[cs], dep[cs]));
                            }
                        }
                        if (t == 0) {
                            double pm, pi, pd; int pe;
                            prev_get(S, PA, id, &pm, &pi, &pd, &pe);
                            acc = xadd(acc, xf(lp.p_DI * lp.p_random * pi, pe));
                        }
                        dv[sl] = acc.v; de[sl] = acc.e; st[sl] = (uint8_t)(stamp0 + t);
                        XF tot = xadd(xf(S.c_dv()[sl], S.c_de()[sl]), acc);
                        S.c_dv()[sl] = tot.v; S.c_de()[sl] = tot.e;
                        fresh = S.d_seen()[sl] ? 0u : 1u;
                    }
                    if (adaptive) {
                        uint32_t n_new;
                        const uint32_t o = n_d + block_prefix(S, fresh, &n_new);
                        if (fresh) { if (o < cap) S.dlist()[o] = (uint16_t)sl; S.d_seen()[sl] = 1; }
                        n_d += n_new;
                    } else if (t == 0 && a < n_l) S.dlist()[a] = (uint16_t)sl;
                }
                if (!adaptive && t == 0) n_d = (uint32_t)n_l;
                __syncthreads();
                src_id = l_id; n_src = n_l;
            }
            if (!ok) { if (tid == 0) { s_fail = SJ_NEED_BIG; s_info = ((unsigned long long)s << 32) | 4u << 16 | n_ent; } __syncthreads(); break; }
            // ---------------- 5b. bm, bi over nodes ; bmb, bib sums (backward.rs:423-555)
            XF pmb = xf_zero(), pib = xf_zero();
            for (uint32_t a = tid; a < n_mi; a += B) {
                uint32_t id = S.act_id()[a], sl = S.act_slot()[a];
                XF am = xf_zero(), ai = xf_zero();
                const uint4 r = __ldg(G.chi_rec + id);
                for (uint32_t k = 0; k < r.y; k++) {
                    const uint32_t ch = k ? G.chi_node[r.x + k] : r.z, eid = k ? G.chi_eid[r.x + k] : r.w;
                    const double tr = trans[eid];
                    double pm, pi, pd; int pe;
                    prev_get(S, PA, ch, &pm, &pi, &pd, &pe);
                    XF tm = xf(tr * (G.emission[ch] == x ? lp.p_match : lp.p_mismatch) * pm, pe);
                    int cs = sp_find(S.ch_key(), S.ch_val(), S.hmask, S.hshift, ch);
                    XF td = cs >= 0 ? xf(tr * S.c_dv()[cs], S.c_de()[cs]) : xf_zero();
                    am = xadd(am, xadd(xmul(tm, lp.p_MM), xmul(td, lp.p_MD)));
                    ai = xadd(ai, xadd(xmul(tm, lp.p_IM), xmul(td, lp.p_ID)));
                }
                double pm, pi, pd; int pe;
                prev_get(S, PA, id, &pm, &pi, &pd, &pe);
                am = xadd(am, xf(lp.p_MI * lp.p_random * pi, pe));
                ai = xadd(ai, xf(lp.p_II * lp.p_random * pi, pe));
                int Em = xexp(am), Ei = xexp(ai), Ec = Em > Ei ? Em : Ei;
                if (Ec == XF_ZERO_E) { S.c_m()[sl] = 0.0; S.c_i()[sl] = 0.0; S.c_mie()[sl] = 0; }
                else { S.c_m()[sl] = am.v == 0.0 ? 0.0 : am.v * pow2i(am.e - Ec); S.c_i()[sl] = ai.v == 0.0 ? 0.0 : ai.v * pow2i(ai.e - Ec); S.c_mie()[sl] = Ec; }
                // begin sums: init_l * (p_XM e_l(x) m''[l] + p_XD d[l])
                XF um = xf((G.emission[id] == x ? lp.p_match : lp.p_mismatch) * pm, pe);
                XF ud = xf(S.c_dv()[sl], S.c_de()[sl]);
                double in = init[id];
                pmb = xadd(pmb, xmul(xadd(xmul(um, lp.p_MM), xmul(ud, lp.p_MD)), in));
                pib = xadd(pib, xmul(xadd(xmul(um, lp.p_IM), xmul(ud, lp.p_ID)), in));
            }
            XF smb = block_xsum(pmb);
            XF sib = block_xsum(pib);
            XF ibn = s_ib;
            __syncthreads();
            XF nmb = xnorm(xadd(smb, xmul(ibn, lp.p_MI * lp.p_random)));
            XF nib = xnorm(xadd(sib, xmul(ibn, lp.p_II * lp.p_random)));
            last_scalar = nmb;
            if (tid == 0) { s_mb = nmb; s_ib = nib; }
        }
        __syncthreads();
        if (n_d > MAX_ACTIVE) { if (tid == 0) s_fail = SJ_CAPACITY; __syncthreads(); break; }

        // ---------------- 7. pack the row into the "previous" arrays, swap hash tables
        for (uint32_t e = tid; e < n_ent; e += B) {
            Cell cl = cell_pack(xf(S.c_m()[e], S.c_mie()[e]), xf(S.c_i()[e], S.c_mie()[e]), xf(S.c_dv()[e], S.c_de()[e]));
            S.p_id()[e] = S.c_id()[e]; S.p_m()[e] = cl.m; S.p_i()[e] = cl.i; S.p_d()[e] = cl.d; S.p_ex()[e] = cl.e;
        }
        S.htog ^= 1u;
        n_prev = n_ent;
        if (tid == 0) s_cells += n_mi;
        __syncthreads();

        // ---------------- 8. store
        if (jb.store) {
            // The page cursor lives in registers, identical in every thread (n_ent and n_d are block-uniform): only a fresh page goes
            // through shared memory, written before the barrier and read after it.  (It used to be a shared variable that thread 0
            // advanced after writing its part of the row with no barrier in between: a slower warp could read the advanced offset and
            // write its entries into the next row's place.)
            const uint64_t bytes = sparse_row_bytes(n_ent, n_d);
            if (bytes > page_left) {
                const uint64_t pg = bytes > SPARSE_PAGE_BYTES ? ((bytes + 255) & ~(uint64_t)255) : SPARSE_PAGE_BYTES;
                if (tid == 0) {
                    unsigned long long o = atomicAdd(io.arena_cursor, (unsigned long long)pg);
                    if (o + pg > io.arena_bytes) s_fail = SJ_OOM;
                    s_page_off = o;
                }
                __syncthreads();
                if (s_fail) break;
                page_off = s_page_off; page_left = pg;
            }
            const uint64_t off = page_off;
            char* pay = io.arena + off;
            double* om = (double*)pay; double* oi = om + n_ent; double* od = oi + n_ent;
            uint32_t* oid = (uint32_t*)(od + n_ent); int* oex = (int*)(oid + n_ent); uint16_t* odl = (uint16_t*)(oex + n_ent);
            for (uint32_t e = tid; e < n_ent; e += B) { om[e] = S.p_m()[e]; oi[e] = S.p_i()[e]; od[e] = S.p_d()[e]; oid[e] = S.p_id()[e]; oex[e] = S.p_ex()[e]; }
            for (uint32_t e = tid; e < n_d; e += B) odl[e] = S.dlist()[e];
            if (tid == 0) {
                RowDesc r;
                r.kind = ROW_SPARSE; r.n_ent = n_ent; r.n_mi = n_mi; r.n_d = n_d; r.off = off;
                if (fwd) { r.mb = xf_zero(); r.ib = s_ib; r.e = last_scalar; }
                else { r.mb = s_mb; r.ib = s_ib; r.e = xf_zero(); }
                io.desc[jb.desc0 + row] = r;
            }
            page_off = off + bytes; page_left -= bytes;
            __syncthreads();
        }
    }
    __syncthreads();
    const bool hand_over = role == 0 && io.rq_ctl && s_fail == SJ_NEED_BIG && n_prev <= io.rq_cap;
    if (hand_over) {   // previous row + scalars to the rescue queue
        char* hr = io.rq_rows + (size_t)job_idx * 32 * io.rq_cap;
        double* hm = (double*)hr; double* hi = hm + io.rq_cap; double* hd = hi + io.rq_cap;
        uint32_t* hid = (uint32_t*)(hd + io.rq_cap); int* hex = (int*)(hid + io.rq_cap);
        for (uint32_t e = tid; e < n_prev; e += B) { hm[e] = S.p_m()[e]; hi[e] = S.p_i()[e]; hd[e] = S.p_d()[e]; hid[e] = S.p_id()[e]; hex[e] = S.p_ex()[e]; }
        if (tid == 0) {
            SHandoff h;
            h.step = (uint32_t)(s_info >> 32); h.n_prev = n_prev; h.mb = s_mb; h.ib = s_ib; h.cells = s_cells; h.last_scalar = last_scalar;
            io.rq_hand[job_idx] = h;
        }
        __threadfence();
        __syncthreads();
        if (tid == 0) {
            const uint32_t slot = atomicAdd(io.rq_ctl, 1u);
            __threadfence();
            *(volatile uint32_t*)(io.rq_items + slot) = job_idx;
        }
    } else if (tid == 0) {
        io.status[job_idx] = s_fail;
        io.final_scalar[job_idx] = last_scalar;
        io.cells[job_idx] = s_fail == SJ_NEED_BIG ? s_info : s_cells;
    }
    if (role == 0 && !io.rq_ctl) return;
    __syncthreads();
  }
}

// ---- gather of the cells of the last dense row that step 0 of a top-n job can read (sparse.h)
__global__ void k_gather_prev0(SGraph G, int dir, uint32_t slot0, const uint32_t* __restrict__ top_ids, const uint32_t* __restrict__ top_cnt,
                               const uint64_t* __restrict__ slabs, const char* __restrict__ pool, uint64_t slab_bytes, uint32_t Np, uint32_t cap,
                               char* __restrict__ out, uint32_t* __restrict__ out_cnt, int* __restrict__ overflow) {
    const uint32_t j = slot0 + blockIdx.x;
    if (slabs[blockIdx.x] == ~0ull) return;
    const char* sl = pool + slabs[blockIdx.x] * slab_bytes;
    const double* gm = (const double*)sl; const double* gi = gm + Np; const double* gd = gi + Np; const int* ge = (const int*)(gd + Np);
    char* o = out + (size_t)j * 32 * cap;
    double* om = (double*)o; double* oi = om + cap; double* od = oi + cap; uint32_t* oid = (uint32_t*)(od + cap); int* oex = (int*)(oid + cap);
    __shared__ uint32_t s_n;
    if (threadIdx.x == 0) s_n = 0;
    __syncthreads();
    auto emit = [&](uint32_t g) {
        const uint32_t k = atomicAdd(&s_n, 1u);
        if (k < cap) { oid[k] = g; om[k] = gm[g]; oi[k] = gi[g]; od[k] = gd[g]; oex[k] = ge[g]; }
    };
    // forward step 0 reads nodes = top + children(top) and their parents (fm, fi) ; backward step 0 reads A0 = top + parents(top) and
    // their children (bd0, bm, bi).  `near` is the direction of the node list, `far` the direction of the reads.
    const uint4* near_rec = dir == 0 ? G.chi_rec : G.par_rec; const uint32_t* near_node = dir == 0 ? G.chi_node : G.par_node;
    const uint4* far_rec = dir == 0 ? G.par_rec : G.chi_rec;  const uint32_t* far_node = dir == 0 ? G.par_node : G.chi_node;
    const uint32_t n_top = top_cnt[j];
    for (uint32_t t = threadIdx.x; t < n_top; t += blockDim.x) {
        const uint32_t id = top_ids[(size_t)j * MAX_ACTIVE + t];
        emit(id);
        const uint4 rf = far_rec[id];
        for (uint32_t a = 0; a < rf.y; a++) emit(a ? far_node[rf.x + a] : rf.z);
        const uint4 rn = near_rec[id];
        for (uint32_t a = 0; a < rn.y; a++) {
            const uint32_t c = a ? near_node[rn.x + a] : rn.z;
            emit(c);
            const uint4 rc = far_rec[c];
            for (uint32_t b = 0; b < rc.y; b++) emit(b ? far_node[rc.x + b] : rc.z);
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) { out_cnt[j] = s_n < cap ? s_n : cap; if (s_n > cap) *overflow = 1; }
}
uint32_t sparse_gather_cap(const dbgphmm_model* m, uint32_t k) {
    const uint64_t D = m->max_deg;
    const uint64_t c = (uint64_t)k * (1 + 2 * D + D * D);
    return (uint32_t)((std::min<uint64_t>(c, 1u << 24) + 1) & ~1ull);
}
int sparse_gather_prev0(dbgphmm_model* m, int dir, uint32_t slot0, uint32_t n, const uint32_t* d_top_ids, const uint32_t* d_top_cnt, const uint64_t* d_slabs,
                        const char* pool, uint64_t slab_bytes, uint32_t Np, uint32_t cap, char* d_out, uint32_t* d_out_cnt, int* d_overflow) {
    if (n == 0) return DBGPHMM_OK;
    SGraph G{m->N, m->E, m->d_emission, m->d_init, m->d_trans, m->d_par_off, m->d_par_node, m->d_par_eid,
             m->d_chi_off, m->d_chi_node, m->d_chi_eid, m->d_pos_of, m->d_par_rec, m->d_chi_rec};
    k_gather_prev0<<<n, 64, 0, MSET(m).stream>>>(G, dir, slot0, d_top_ids, d_top_cnt, d_slabs, pool, slab_bytes, Np, cap, d_out, d_out_cnt, d_overflow);
    COUNT_LAUNCH();
    CUDA_TRY(cudaGetLastError());
    return DBGPHMM_OK;
}

static size_t sparse_smem_bytes(uint32_t cap, uint32_t hcap) { return SS_BYTES(cap, hcap); }

int sparse_configure(dbgphmm_model* m) {
    (void)m;
    CUDA_TRY(cudaFuncSetAttribute(k_sparse<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    CUDA_TRY(cudaFuncSetAttribute(k_sparse<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    return DBGPHMM_OK;
}

// threads per job: the rows hold 40-130 entries and every phase is a short dependent chain, so two warps do as well as
// four and let more jobs stay resident
static int sparse_threads(uint32_t cap) { return cap <= 64 ? 32 : (cap <= 256 ? 64 : 256); }

uint32_t sparse_default_cap() {
    if (const char* e = getenv("DBGPHMM_SPARSE_CAP")) { int c0 = atoi(e); if (c0 >= 32 && c0 <= 832) return ((uint32_t)c0 + 15u) & ~15u; }
    return 128;   // rows of a top-n (n_active_nodes = 40) job hold 50-110 entries ; the rescue launch takes the exceptions
}

uint32_t sparse_rescue_cap(uint32_t cap) {
    // 192 entries: with the 128-entry primary tables that leaves shared memory for 9 primary jobs per SM beside one rescue CTA (256 -> 8).
    // No row of the C3 / C5 workloads outgrows it (a row that does is re-run with 256 / 832 entries: correct, one job latency slower).
    uint32_t rc = 192;
    if (const char* e = getenv("DBGPHMM_SPARSE_RESCUE")) { const int v = atoi(e); if (v <= 0) return 0; if (v >= 64 && v <= 832) rc = ((uint32_t)v + 15u) & ~15u; }
    return cap < rc ? rc : 0;
}

static uint32_t hcap_of(uint32_t cap) {
    uint32_t hcap = 1;
    while (hcap < 2 * cap) hcap <<= 1;
    return hcap;
}

// jobs resident at once: as many primary CTAs per SM as shared memory allows beside one rescue CTA
uint32_t sparse_wave_jobs(dbgphmm_model* m, uint32_t cap) {
    const uint32_t rcap = sparse_rescue_cap(cap);
    int per_sm = 0;
    cudaFuncSetAttribute(k_sparse<0>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    cudaFuncSetAttribute(k_sparse<1>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_sparse<1>, sparse_threads(cap), sparse_smem_bytes(cap, hcap_of(cap))) != cudaSuccess || per_sm < 1) {
        cudaGetLastError();
        per_sm = 1;
    }
    if (rcap) {
        const size_t margin = 4096;   // bytes kept free per SM beside the primary CTAs and the rescue CTA
        const size_t sm_bytes = 227 * 1024, one = sparse_smem_bytes(cap, hcap_of(cap)) + 1024, big = sparse_smem_bytes(rcap, hcap_of(rcap)) + 1024;
        const int fit = (int)((sm_bytes - big - margin) / one);
        if (fit >= 1 && fit < per_sm) per_sm = fit;
    }
    return (uint32_t)per_sm * (uint32_t)m->n_sm;
}

// Both directions' sparse launches (primary + rescue each) resident at once: the two-thread bulk path runs the two sparse phases side by
// side only then.  With more jobs one direction's rescue CTAs would find no shared memory until primary CTAs of the other direction leave
// -- persistent ones, i.e. at the end -- and the jobs handed over to them would start a whole job latency late.
bool sparse_pair_fits(dbgphmm_model* m, uint32_t cap, uint32_t n_jobs) {
    const uint32_t rcap = sparse_rescue_cap(cap);
    const size_t sm_bytes = 227 * 1024, one = sparse_smem_bytes(cap, hcap_of(cap)) + 1024, big = rcap ? sparse_smem_bytes(rcap, hcap_of(rcap)) + 1024 : 0;
    // on average over the SMs: the rescue CTAs are a pool (any of them takes any handed-over job), so it is enough that most are resident
    const double per_sm = (double)n_jobs / (double)m->n_sm;
    return 2.0 * per_sm * (double)one + 2.0 * (double)big + 4096.0 <= (double)sm_bytes;
}

static int sparse_launch(dbgphmm_model* m, cudaStream_t st, uint32_t grid, const SGraph& G, const SJob* d_jobs, const SparseIO& io, uint32_t cap, int role, int dir) {
    const uint32_t hcap = hcap_of(cap);
    const size_t smem = sparse_smem_bytes(cap, hcap);
    if (smem > 200 * 1024) { dbg_set_error("sparse_run: capacity too large for shared memory"); return DBGPHMM_ERR_INVALID; }
    const int threads = sparse_threads(cap);
    if (dir == 0) k_sparse<0><<<grid, threads, smem, st>>>(G, m->lin, d_jobs, io, cap, hcap, role);
    else k_sparse<1><<<grid, threads, smem, st>>>(G, m->lin, d_jobs, io, cap, hcap, role);
    COUNT_LAUNCH();
    CUDA_TRY(cudaGetLastError());
    return DBGPHMM_OK;
}

int sparse_run(dbgphmm_model* m, const SJob* d_jobs, uint32_t n_jobs, const SparseIO& io_in, uint32_t cap, int dir, uint32_t rescue_cap) {
    if (n_jobs == 0) return DBGPHMM_OK;
    SGraph G{m->N, m->E, m->d_emission, m->d_init, m->d_trans, m->d_par_off, m->d_par_node, m->d_par_eid,
             m->d_chi_off, m->d_chi_node, m->d_chi_eid, m->d_pos_of, m->d_par_rec, m->d_chi_rec};
    // the kernel is latency-bound: as many resident jobs per SM as shared memory allows
    cudaFuncSetAttribute(k_sparse<0>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    cudaFuncSetAttribute(k_sparse<1>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    SparseIO io = io_in;
    if (!rescue_cap || !io.rq_ctl) {
        io.rq_ctl = nullptr;
        return sparse_launch(m, MSET(m).stream, n_jobs, G, d_jobs, io, cap, 0, dir);
    }
    // primary: as many persistent CTAs as stay resident beside one rescue CTA per SM ; rescue CTAs idle until a job is handed over
    const uint32_t grid = std::min<uint32_t>(n_jobs, sparse_wave_jobs(m, cap));
    io.rq_cap = cap; io.rq_n_primary = grid; io.rq_n_jobs = n_jobs;
    CUDA_TRY(cudaEventRecord(MSET(m).ev_fork, MSET(m).stream));
    CUDA_TRY(cudaStreamWaitEvent(MSET(m).aux, MSET(m).ev_fork, 0));
    ST_TRY(sparse_launch(m, MSET(m).stream, grid, G, d_jobs, io, cap, 0, dir));
    ST_TRY(sparse_launch(m, MSET(m).aux, std::min<uint32_t>(n_jobs, (uint32_t)m->n_sm), G, d_jobs, io, rescue_cap, 1, dir));
    CUDA_TRY(cudaEventRecord(MSET(m).ev_join, MSET(m).aux));
    CUDA_TRY(cudaStreamWaitEvent(MSET(m).stream, MSET(m).ev_join, 0));
    return DBGPHMM_OK;
}
