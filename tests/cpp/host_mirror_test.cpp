// host_mirror_test.cpp -- TEST INFRASTRUCTURE.  Exercises include/zebra_b200.hpp (the C++ host mirror of the reference's
// LSHIndex / Database / metric interface, /root/reference/src/database/index/lsh.rs:145-566, src/database/core.rs:55-381,
// src/distance.rs) against the CPU oracle (oracle/zb_oracle.c, linked as libzb_oracle.so).
//   host_mirror_test cpu <dir>   no device needed: types, error behaviour, the store dump written for the Python side to read
//   host_mirror_test read <file> reads a store dump written by zebra_b200/interchange.py and prints its summary
//   host_mirror_test gpu <dir>   parity on a B200: ids / distance bits bit-exact against the oracle through the C++ API
#include <cstdio>
#include <cstdlib>
#include <string>

#include "../../include/zebra_b200.hpp"

extern "C" {  // the oracle's C entry points (oracle/zb_oracle.c)
struct zbo_index;
zbo_index* zbo_create(int dim, int metric, uint64_t max_node_size, int num_trees, uint64_t seed);
void zbo_destroy(zbo_index* ix);
int zbo_add(zbo_index* ix, uint64_t n, const float* rows, uint64_t* out_ids);
int zbo_remove(zbo_index* ix, uint64_t n, const uint64_t* ids, uint8_t* out_removed);
int64_t zbo_deduplicate(zbo_index* ix, uint64_t* out_ids, uint64_t cap);
int zbo_search_batch(const zbo_index* ix, uint64_t nq, const float* queries, uint64_t top_k, int nthreads, uint64_t* out_ids,
                     uint64_t* out_bits, uint32_t* out_counts);
uint64_t zbo_distance_bits(int metric, const float* row, const float* query, int n);
int zbo_point_is_above(const float* coef, float constant, const float* x, int n);
}

using namespace zebra;
static int failures = 0;
#define EXPECT(cond)                                                        \
    do {                                                                    \
        if (!(cond)) {                                                      \
            std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond);   \
            ++failures;                                                     \
        }                                                                   \
    } while (0)

constexpr size_t N = 64;
static uint64_t lcg_state = 12345;
static float frand() {  // uniform in [-1, 1), exact in f32
    lcg_state = lcg_state * 6364136223846793005ull + 1442695040888963407ull;
    return (float)((int32_t)(lcg_state >> 32) >> 8) * (1.0f / 8388608.0f);
}
static std::vector<Embedding<N>> clustered(size_t n, size_t centres) {
    std::vector<Embedding<N>> c(centres), out(n);
    for (auto& v : c)
        for (auto& x : v) x = 2.0f * frand();
    for (size_t i = 0; i < n; ++i)
        for (size_t k = 0; k < N; ++k) out[i][k] = c[i % centres][k] + 0.25f * frand();
    return out;
}

static int run_cpu(const std::string& dir) {
    static_assert(sizeof(Embedding<384>) == 1536 && sizeof(Embedding<768>) == 3072, "Embedding<N> layout (lib.rs:16-18)");
    static_assert(CosineDistance<8>::ZB_METRIC == 0 && L2SquaredDistance<8>::ZB_METRIC == 1 && L2Distance<8>::ZB_METRIC == 2, "");
    static_assert(HammingDistance<8>::ZB_METRIC == 9 && PNormDistance<8>::ZB_METRIC == 11 && MinkowskiDistance<8>::HAS_POWER, "");
    EXPECT((LSHIndexOptions<N>{}.max_node_size == 5 && LSHIndexOptions<N>{}.num_trees == 15));  // lsh.rs:131-138
    EXPECT(MinkowskiDistance<N>{}.power == 0);                                                  // #[derive(Default)]
    Embedding<N> zero;
    EXPECT(zero[0] == 0.0f && zero[N - 1] == 0.0f);                                             // lib.rs:31-35
    bool threw = false;
    try { Embedding<N>::try_from(std::vector<float>(N + 1)); } catch (const Error& e) { threw = e.code == ZB_ERR_INVALID; }
    EXPECT(threw);
    Uuid a = Uuid::now_v7(), b = Uuid::now_v7();
    EXPECT(a.version() == 7 && (a.bytes[8] >> 6) == 2 && a != b && a.simple().size() == 32);
    Uuid lo, hi;
    lo.bytes[15] = 1;
    hi.bytes[0] = 1;
    EXPECT(lo < hi && !(hi < lo));                                                              // big-endian order
    int ndev = 0;
    zb_device_count(&ndev);
    if (ndev == 0) {  // no CPU fallback anywhere: constructors and metrics throw
        int code = 0;
        try { LSHIndex<N>::create(a, LSHIndexOptions<N>{}, L2SquaredDistance<N>{}); } catch (const Error& e) { code = e.code; }
        EXPECT(code == ZB_ERR_NO_DEVICE);
        code = 0;
        try { CosineDistance<N>{}.distance(zero, zero); } catch (const Error& e) { code = e.code; }
        EXPECT(code == ZB_ERR_NO_DEVICE);
        code = 0;
        try { Database<N, CosineDistance<N>>::new_with_path(dir + "/nodev.zebra", LSHIndexOptions<N>{}); } catch (const Error& e) { code = e.code; }
        EXPECT(code == ZB_ERR_NO_DEVICE);
    }
    // a store dump for the Python side (tests/test_cpp_host.py reads it with zebra_b200.interchange.read_store)
    store::Dump d;
    d.dim = 4;
    uint8_t zb[64];
    uint64_t need = 0;
    Uuid dbid;
    for (int i = 0; i < 16; ++i) dbid.bytes[i] = (uint8_t)(0x10 + i);
    check(zb_zebra_file_encode(dbid.bytes.data(), ZB_METRIC_MINKOWSKI, 3, 5, 2, zb, sizeof zb, &need));
    EXPECT(need == 44);
    d.zebra.assign(zb, zb + need);
    for (int i = 0; i < 3; ++i) {
        Uuid u;
        u.bytes[15] = (uint8_t)(i + 1);
        d.ids.push_back(u);
        for (int k = 0; k < 4; ++k) d.rows.push_back((float)(10 * i + k) * 0.5f);
    }
    // tree 0: Inner(plane (1,0,0,-1), 0.25, Leaf[id1], Leaf[id2, id3]); tree 1: root leaf of all three
    const int32_t nodes[] = {0, 1, 2, -1, -1, -1, -1, 0, -1, -1, -1, 1, -1, -1, -1, 2};
    const float coef[] = {1.f, 0.f, 0.f, -1.f}, cst[] = {0.25f};
    const int64_t leaf_off[] = {0, 1, 3, 6};
    uint8_t mids[6 * 16];
    const int who[] = {0, 1, 2, 0, 1, 2};
    for (int j = 0; j < 6; ++j) std::memcpy(mids + 16 * j, d.ids[who[j]].bytes.data(), 16);
    for (int t = 0; t < 2; ++t) {
        Bytes blob(4096);
        check(zb_tree_blob_encode(4, 4, nodes, t == 0 ? 0 : 3, coef, cst, leaf_off, mids, blob.data(), blob.size(), &need));
        blob.resize(need);
        d.trees.emplace_back(store::tree_key((uint32_t)t), blob);
    }
    EXPECT(d.trees[0].second.size() == 4 + 16 + 4 + (4 + 8 + 24) + (4 + 8 + 48) && d.trees[1].second.size() == 4 + 8 + 72);
    store::write(dir + "/cpp.store", d);
    store::Dump r = store::read(dir + "/cpp.store");
    EXPECT(r.dim == 4 && r.zebra == d.zebra && r.trees == d.trees && r.ids == d.ids && r.rows == d.rows);
    threw = false;
    try { store::read(dir + "/missing.store"); } catch (const Error&) { threw = true; }
    EXPECT(threw);
    return failures;
}

static int run_read(const std::string& file) {
    store::Dump d = store::read(file);
    std::printf("dim=%u zebra=%zu trees=%zu rows=%zu first_row0=%g last_id_byte=%u tree0_bytes=%zu\n", d.dim, d.zebra.size(), d.trees.size(),
                d.ids.size(), d.rows.empty() ? 0.0 : (double)d.rows[0], d.ids.empty() ? 0u : (unsigned)d.ids.back().bytes[15],
                d.trees.empty() ? (size_t)0 : d.trees[0].second.size());
    return 0;
}

template <class Met>
static void compare_search(const char* tag, const LSHIndex<N>& ix, zbo_index* orc, const std::map<Uuid, uint64_t>& ordinal_of,
                           const std::vector<Embedding<N>>& queries, size_t k, const Met& metric) {
    auto got = ix.search_batch(queries, k);
    std::vector<uint64_t> eo(queries.size() * k), eb(queries.size() * k);
    std::vector<uint32_t> ec(queries.size());
    zbo_search_batch(orc, queries.size(), queries[0].data(), k, 4, eo.data(), eb.data(), ec.data());
    bool ok = true;
    for (size_t q = 0; q < queries.size() && ok; ++q) {
        ok = got[q].size() == ec[q];
        for (size_t i = 0; ok && i < ec[q]; ++i)
            ok = ordinal_of.at(got[q][i].first) == eo[q * k + i] && got[q][i].second == eb[q * k + i];
    }
    auto one = ix.search(queries[3], k, metric);  // the single-query form of lsh.rs:544
    ok = ok && one == got[3];
    std::printf("[cpp host] %s: %s\n", tag, ok ? "ok" : "MISMATCH");
    if (!ok) ++failures;
}

static int run_gpu(const std::string& dir) {
    const size_t n = 3000, nq = 96, k = 10;
    auto rows = clustered(n, 16);
    std::vector<Embedding<N>> queries(rows.begin(), rows.begin() + nq / 2);
    for (size_t i = 0; i < nq / 2; ++i) {
        Embedding<N> q;
        for (auto& x : q) x = 2.0f * frand();
        queries.push_back(q);
    }
    // ---- LSHIndex: add / search / remove / deduplicate against the oracle (reference defaults 5 / 15) ----
    {
        L2SquaredDistance<N> metric;
        LSHIndex<N> ix = LSHIndex<N>::create(Uuid::now_v7(), LSHIndexOptions<N>{}, metric, 0, 7);
        EXPECT(ix.is_empty() && ix.no_vectors() && ix.no_trees());
        zbo_index* orc = zbo_create((int)N, 1, 5, 15, 7);
        std::vector<Uuid> ids = ix.add(rows);
        std::vector<uint64_t> oids(n);
        zbo_add(orc, n, rows[0].data(), oids.data());
        std::map<Uuid, uint64_t> ordinal_of;
        for (size_t i = 0; i < n; ++i) ordinal_of[ids[i]] = i;
        EXPECT(ordinal_of.size() == n && ids[0] < ids[1] && ids[0].version() == 7 && !ix.is_empty());
        compare_search("LSHIndex add + search_batch", ix, orc, ordinal_of, queries, k, metric);
        std::vector<Uuid> dead;
        std::vector<uint64_t> dead_o;
        for (size_t i = 0; i < n; i += 7) { dead.push_back(ids[i]); dead_o.push_back(i); }
        dead.push_back(Uuid::now_v7());  // never inserted
        std::set<Uuid> removed = ix.remove(dead);
        EXPECT(removed.size() == dead_o.size() && !removed.count(dead.back()));
        zbo_remove(orc, dead_o.size(), dead_o.data(), nullptr);
        compare_search("after remove", ix, orc, ordinal_of, queries, k, metric);
        bool threw = false;
        try { ix.search(queries[0], k, CosineDistance<N>{}); } catch (const Error& e) { threw = e.code == ZB_ERR_INVALID; }
        EXPECT(threw);
        std::vector<Embedding<N>> dups(rows.begin() + 1, rows.begin() + 41);  // copies of live rows (1..40 minus the removed 7, 14, ...)
        std::vector<Uuid> dup_ids = ix.add(dups);
        std::vector<uint64_t> dup_o(dups.size());
        zbo_add(orc, dups.size(), dups[0].data(), dup_o.data());
        for (size_t i = 0; i < dups.size(); ++i) ordinal_of[dup_ids[i]] = dup_o[i];
        std::set<Uuid> gone = ix.deduplicate();
        std::vector<uint64_t> exp(4096);
        const int64_t cnt = zbo_deduplicate(orc, exp.data(), exp.size());
        bool same = (int64_t)gone.size() == cnt && cnt > 0;
        for (int64_t i = 0; same && i < cnt; ++i) {
            bool found = false;
            for (const Uuid& u : gone) found = found || ordinal_of.at(u) == exp[i];
            same = found;
        }
        std::printf("[cpp host] deduplicate removed %zu: %s\n", gone.size(), same ? "ok" : "MISMATCH");
        if (!same) ++failures;
        compare_search("after deduplicate", ix, orc, ordinal_of, queries, k, metric);
        std::vector<uint64_t> keys;
        std::vector<uint32_t> depths;
        ix.hash(queries, keys, depths);
        EXPECT(keys.size() == nq * 15 && depths[0] > 0);
        ix.clear();
        EXPECT(ix.is_empty());
        zbo_destroy(orc);
    }
    // ---- the Metric trait and point_is_above for every metric struct ----
    {
        bool ok = true;
        const Embedding<N>&a = rows[1], &b = queries[nq - 1];
        ok = ok && CosineDistance<N>{}.distance(a, b) == zbo_distance_bits(0, a.data(), b.data(), N);
        ok = ok && L2SquaredDistance<N>{}.distance(a, b) == zbo_distance_bits(1, a.data(), b.data(), N);
        ok = ok && L2Distance<N>{}.distance(a, b) == zbo_distance_bits(2, a.data(), b.data(), N);
        ok = ok && ChebyshevDistance<N>{}.distance(a, b) == zbo_distance_bits(3, a.data(), b.data(), N);
        ok = ok && CanberraDistance<N>{}.distance(a, b) == zbo_distance_bits(4, a.data(), b.data(), N);
        ok = ok && BrayCurtisDistance<N>{}.distance(a, b) == zbo_distance_bits(5, a.data(), b.data(), N);
        ok = ok && ManhattanDistance<N>{}.distance(a, b) == zbo_distance_bits(6, a.data(), b.data(), N);
        ok = ok && L3Distance<N>{}.distance(a, b) == zbo_distance_bits(7, a.data(), b.data(), N);
        ok = ok && L4Distance<N>{}.distance(a, b) == zbo_distance_bits(8, a.data(), b.data(), N);
        ok = ok && HammingDistance<N>{}.distance(a, b) == zbo_distance_bits(9, a.data(), b.data(), N);
        MinkowskiDistance<N> mk;
        mk.power = 3;
        ok = ok && mk.distance(a, b) == zbo_distance_bits(10 | (3 << 8), a.data(), b.data(), N);
        PNormDistance<N> pn;
        pn.power = 2;
        ok = ok && pn.distance(a, b) == zbo_distance_bits(11 | (2 << 8), a.data(), b.data(), N);
        ok = ok && MinkowskiDistance<N>{}.distance(a, b) == 0x7F800000ull;  // Default power 0: powf(N, +inf)
        Hyperplane<N> h;
        h.coefficients = rows[2];
        h.constant = -0.5f;
        for (size_t i = 0; i < 32; ++i)
            ok = ok && h.point_is_above(rows[i]) == (zbo_point_is_above(h.coefficients.data(), h.constant, rows[i].data(), N) != 0);
        std::printf("[cpp host] Metric::distance x 13 + point_is_above: %s\n", ok ? "ok" : "MISMATCH");
        if (!ok) ++failures;
    }
    // ---- Database: insert_records / query_vectors / remove / save_database / open with a scalar metric ----
    {
        using Db = Database<N, ManhattanDistance<N>>;
        LSHIndexOptions<N> opt;
        opt.max_node_size = 32;
        opt.num_trees = 4;
        const std::string path = dir + "/cpp_db.zebra";
        Db db = Db::new_with_path(path, opt);
        EXPECT(db.query_vectors(queries, 5).empty());                   // core.rs:294: no vectors -> empty map
        std::vector<Bytes> docs(n);
        for (size_t i = 0; i < n; ++i) docs[i] = Bytes{(uint8_t)i, (uint8_t)(i >> 8)};
        std::vector<Uuid> ids = db.insert_records(rows, docs);
        // the device build is seeded by the index (seed 0 here); the oracle with the same seed builds the same forest
        zbo_index* orc = zbo_create((int)N, 6, 32, 4, 0);
        std::vector<uint64_t> oids(n);
        zbo_add(orc, n, rows[0].data(), oids.data());
        auto res = db.query_vectors(queries, k);
        std::vector<uint64_t> eo(nq * k), eb(nq * k);
        std::vector<uint32_t> ec(nq);
        zbo_search_batch(orc, nq, queries[0].data(), k, 4, eo.data(), eb.data(), ec.data());
        bool ok = res.size() == nq;
        for (size_t q = 0; ok && q < nq; ++q) {
            std::set<size_t> want, have;
            for (uint32_t i = 0; i < ec[q]; ++i) want.insert((size_t)eo[q * k + i]);
            for (const auto& kv : res[q]) have.insert((size_t)kv.second[0] | ((size_t)kv.second[1] << 8));
            ok = want == have;
        }
        std::printf("[cpp host] Database insert_records + query_vectors (Manhattan): %s\n", ok ? "ok" : "MISMATCH");
        if (!ok) ++failures;
        db.remove({ids[0], ids[1]});
        auto after = db.query_vectors({rows[0]}, k);
        EXPECT(!after[0].count(ids[0]) && !after[0].count(ids[1]));
        db.save_database();
        Db db2 = Db::open(path);
        EXPECT(db2.uuid() == db.uuid() && db2.index_options() == opt);
        auto s1 = db.index.search_batch(queries, k), s2 = db2.index.search_batch(queries, k);
        std::printf("[cpp host] Database save_database + open: %s\n", s1 == s2 ? "ok" : "MISMATCH");
        if (!(s1 == s2)) ++failures;
        EXPECT(db2.index.stats().live_rows == n - 2);
        Db db3 = Db::open_or_create(dir + "/fresh.zebra", opt);          // nothing to open: created
        EXPECT(db3.index.is_empty());
        db3.clear_database();
        db2.clear_database();
        EXPECT(!std::ifstream(path).good());
        zbo_destroy(orc);
    }
    return failures;
}

int main(int argc, char** argv) {
    if (argc < 3) {
        std::printf("usage: host_mirror_test cpu|gpu <dir> | read <file>\n");
        return 2;
    }
    const std::string mode = argv[1], arg = argv[2];
    int rc = 0;
    try {
        rc = mode == "cpu" ? run_cpu(arg) : (mode == "read" ? run_read(arg) : run_gpu(arg));
    } catch (const zebra::Error& e) {
        std::printf("zebra::Error %d: %s\n", e.code, e.what());
        rc = 1;
    }
    std::printf("%s\n", rc ? "FAILED" : "all ok");
    return rc ? 1 : 0;
}
