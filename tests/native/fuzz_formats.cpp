// Sanitizer fuzz driver of the host-only text readers (csrc/formats.cu: DBG, MAP and dataset-JSON files; csrc/score.cu: the
// Euler-circuit count).  tests/test_formats.py compiles those two sources as plain C++ with -fsanitize=address,undefined together with
// this file and runs it: seeded random mutations of valid texts are parsed; whatever parses is walked, written back and parsed again.
// The readers must either accept or return a status -- never read or write out of bounds, overflow a signed integer or leak.
// The entry points of the device side that formats.cu refers to are stubbed (no CUDA in this program).
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <string>
#include <vector>
#include "../../dbgphmm_b200/csrc/model.h"
#include "../../include/dbgphmm_b200.h"

static std::string g_error;
void dbg_set_error(const std::string& s) { g_error = s; }
extern "C" const char* dbgphmm_last_error(void) { return g_error.c_str(); }
extern "C" int dbgphmm_model_create(uint32_t, uint32_t, const uint32_t*, const uint32_t*, const uint8_t*, const double*, const double*, const dbgphmm_params*, int, uint64_t,
                                    dbgphmm_model**) { return DBGPHMM_ERR_CUDA; }
extern "C" void dbgphmm_model_destroy(dbgphmm_model*) {}
extern "C" int dbgphmm_model_set_copy_nums_batch(dbgphmm_model*, uint32_t, const uint32_t*, int) { return DBGPHMM_ERR_CUDA; }
extern "C" int dbgphmm_reads_create(uint64_t n, const uint64_t* off, const uint8_t* bases, dbgphmm_reads** out) {
    dbgphmm_reads* r = new dbgphmm_reads();
    r->n_reads = n; r->off.assign(off, off + n + 1); r->bases.assign(bases, bases + off[n]);
    *out = r;
    return DBGPHMM_OK;
}

static std::string slurp(const char* path) {
    std::string s; FILE* f = fopen(path, "rb");
    if (!f) { fprintf(stderr, "cannot open %s\n", path); exit(2); }
    char buf[65536]; size_t n;
    while ((n = fread(buf, 1, sizeof buf, f)) > 0) s.append(buf, n);
    fclose(f);
    return s;
}

static std::string mutate(const std::string& base, std::mt19937_64& rng, const char* alphabet) {
    std::string b = base;
    const size_t na = strlen(alphabet);
    const int n_ops = 1 + (int)(rng() % 6);
    for (int o = 0; o < n_ops; o++) {
        if (b.empty()) { b = "{"; continue; }
        const size_t pos = rng() % b.size();
        switch (rng() % 5) {
            case 0: b[pos] = alphabet[rng() % na]; break;
            case 1: b.erase(pos, 1 + rng() % 6); break;
            case 2: { std::string ins; for (int i = 0, n = 1 + (int)(rng() % 4); i < n; i++) ins += alphabet[rng() % na]; b.insert(pos, ins); break; }
            case 3: b.resize(pos); break;                                                  // truncation
            default: { const size_t len = 1 + rng() % 24; b.insert(pos, b.substr(pos, len)); break; }   // a repeated piece
        }
    }
    return b;
}

template <typename F> static std::string text_of(F&& to_text) {
    uint64_t need = 0;
    if (to_text(nullptr, 0, &need) != DBGPHMM_OK) return std::string();
    std::string s(need, '\0');
    if (to_text(&s[0], need, &need) != DBGPHMM_OK) { fprintf(stderr, "to_text failed on the second call: %s\n", dbgphmm_last_error()); exit(1); }
    return s;
}

int main(int argc, char** argv) {
    if (argc != 5) { fprintf(stderr, "usage: fuzz_formats <dbg text> <map text> <dataset json> <iterations>\n"); return 2; }
    const std::string seeds[3] = {slurp(argv[1]), slurp(argv[2]), slurp(argv[3])};
    const long iters = atol(argv[4]);
    std::mt19937_64 rng(20251019);
    long parsed[3] = {0, 0, 0}, rejected[3] = {0, 0, 0};
    for (long it = 0; it < iters; it++) {
        const int kind = (int)(it % 3);
        const std::string t = it < 3 ? seeds[kind] : mutate(seeds[kind], rng, kind == 2 ? "0123456789,:\"{}[]-+.eEILCACGTn( )\\u" : "0123456789,\t\n:-KNEnACGT. ");
        if (kind == 0) {
            dbgphmm_dbg* d = nullptr;
            if (dbgphmm_dbg_from_text(t.data(), t.size(), &d) != DBGPHMM_OK) { rejected[0]++; if (it < 3) { fprintf(stderr, "seed DBG rejected: %s\n", dbgphmm_last_error()); return 1; } continue; }
            uint32_t sz[6];
            dbgphmm_dbg_sizes(d, sz);
            std::vector<uint32_t> es(sz[5]), ed(sz[5]), cn(sz[2]), ce(sz[2]), cc(sz[4]), full(sz[2]); std::vector<uint8_t> em(sz[2]);   // PHMM nodes = full edges
            dbgphmm_dbg_phmm_graph(d, es.data(), ed.data(), em.data(), cn.data(), ce.data());
            dbgphmm_dbg_get_copy_nums(d, cc.data());
            dbgphmm_dbg_expand_copy_nums(d, 1, cc.data(), full.data());
            uint64_t gs = 0; double ne = 0.0;
            dbgphmm_dbg_genome_size(d, 1, cc.data(), &gs);
            dbgphmm_dbg_n_euler_circuits(d, 1, cc.data(), &ne);       // may fail: copy numbers that no longer balance
            dbgphmm_dbg_set_copy_nums(d, cc.data());
            const std::string back = text_of([&](char* b, uint64_t c, uint64_t* n) { return dbgphmm_dbg_to_text(d, b, c, n); });
            dbgphmm_dbg* d2 = nullptr;
            if (dbgphmm_dbg_from_text(back.data(), back.size(), &d2) != DBGPHMM_OK) { fprintf(stderr, "a written DBG does not parse: %s\n%s\n", dbgphmm_last_error(), back.c_str()); return 1; }
            const std::string back2 = text_of([&](char* b, uint64_t c, uint64_t* n) { return dbgphmm_dbg_to_text(d2, b, c, n); });
            if (back2 != back) { fprintf(stderr, "DBG text is not a fixed point\n"); return 1; }
            dbgphmm_dbg_destroy(d2); dbgphmm_dbg_destroy(d);
            parsed[0]++;
        } else if (kind == 1) {
            dbgphmm_mappings* mp = nullptr;
            if (dbgphmm_mappings_from_map_text(t.data(), t.size(), &mp) != DBGPHMM_OK) { rejected[1]++; if (it < 3) { fprintf(stderr, "seed MAP rejected: %s\n", dbgphmm_last_error()); return 1; } continue; }
            // (the accessors of the handle live with the device code: the CSR is checked in place)
            bool ok = !mp->read_off.empty() && mp->read_off[0] == 0 && mp->read_off.back() + 1 == mp->row_off.size() && mp->row_off[0] == 0 &&
                      mp->row_off.back() == mp->nodes.size() && mp->nodes.size() == mp->logp.size();
            for (size_t i = 1; ok && i < mp->read_off.size(); i++) ok = mp->read_off[i - 1] <= mp->read_off[i];
            for (size_t i = 1; ok && i < mp->row_off.size(); i++) ok = mp->row_off[i - 1] <= mp->row_off[i];
            if (!ok) { fprintf(stderr, "MAP reader built an inconsistent CSR from:\n%s\n", t.c_str()); return 1; }
            delete mp;
            parsed[1]++;
        } else {
            dbgphmm_dataset* ds = nullptr;
            if (dbgphmm_dataset_from_json_text(t.data(), t.size(), &ds) != DBGPHMM_OK) { rejected[2]++; if (it < 3) { fprintf(stderr, "seed dataset rejected: %s\n", dbgphmm_last_error()); return 1; } continue; }
            uint64_t sz[5];
            dbgphmm_dataset_sizes(ds, sz);
            std::vector<uint64_t> ho(sz[0] + 1), ro(sz[2] + 1), op(sz[3]); std::vector<uint8_t> hb(sz[1]), hs(sz[0]), rb(sz[3]), rc(sz[2]); std::vector<int64_t> oh(sz[3]);
            dbgphmm_dataset_genome(ds, ho.data(), hb.data(), hs.data());
            dbgphmm_dataset_read_origins(ds, ro.data(), rb.data(), rc.data(), oh.data(), op.data());
            dbgphmm_params par; dbgphmm_dataset_params(ds, &par);
            dbgphmm_reads* rd = nullptr;
            if (dbgphmm_dataset_reads(ds, &rd) == DBGPHMM_OK) delete rd;
            const std::string back = text_of([&](char* b, uint64_t c, uint64_t* n) { return dbgphmm_dataset_to_json_text(ds, b, c, n); });
            dbgphmm_dataset* d2 = nullptr;
            if (dbgphmm_dataset_from_json_text(back.data(), back.size(), &d2) != DBGPHMM_OK) { fprintf(stderr, "a written dataset does not parse: %s\n%s\n", dbgphmm_last_error(), back.c_str()); return 1; }
            const std::string back2 = text_of([&](char* b, uint64_t c, uint64_t* n) { return dbgphmm_dataset_to_json_text(d2, b, c, n); });
            if (back2 != back) { fprintf(stderr, "dataset JSON is not a fixed point:\n%s\n%s\n", back.c_str(), back2.c_str()); return 1; }
            dbgphmm_dataset_destroy(d2); dbgphmm_dataset_destroy(ds);
            parsed[2]++;
        }
    }
    printf("dbg parsed %ld rejected %ld; map parsed %ld rejected %ld; dataset parsed %ld rejected %ld\n", parsed[0], rejected[0], parsed[1], rejected[1], parsed[2], rejected[2]);
    return 0;
}
