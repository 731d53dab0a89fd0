// fuzz_interchange.cpp -- TEST INFRASTRUCTURE.  Mutation fuzzing of the stored-value parsers of libzebra_b200
// (zb_tree_blob_decode, zb_store_flatten, zb_zebra_file_decode in zebra_b200/csrc/zb_interchange.cpp): they take bytes an
// on-disk store hands them (/root/reference/src/database/index/lsh.rs:99-119), so truncated, bit-flipped or spliced input
// must come back as ZB_OK or ZB_ERR_INVALID -- never a crash, an out-of-bounds access or an unbounded allocation.
// tests/test_interchange.py compiles THIS file together with zb_interchange.cpp under -fsanitize=address,undefined (the
// device entry points the importer would call are stubbed out below) and runs it for a bounded number of iterations.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/zebra_b200.h"

// ---- stubs: what zb_interchange.cpp needs from the rest of the library (no device in this build) ----
namespace zb {
thread_local std::string t_last_error;
size_t g_device_bytes = 0;
}
// A canned 2-tree forest over 5 rows (row 3 removed) behind the handle FAKE, so the export entry points can be driven on the
// CPU: tree 0 = Inner(plane 0, Leaf[0,1], Inner(plane 1, Leaf[2,3], Leaf[4])), tree 1 = Leaf[0..4].  dim = 2.
static zb_index* const FAKE = reinterpret_cast<zb_index*>(0x1);
static const int32_t F_NODES[] = {0, 1, 2, -1, -1, -1, -1, 0, 1, 3, 4, -1, -1, -1, -1, 1, -1, -1, -1, 2, -1, -1, -1, 3};
static const int32_t F_ROOTS[] = {0, 5};
static const float F_COEF[] = {1.f, -1.f, 0.5f, 2.f}, F_CST[] = {0.25f, -3.f};
static const int64_t F_LEAF_OFF[] = {0, 2, 4, 5, 10};
static const uint64_t F_MEMBERS[] = {0, 1, 2, 3, 4, 0, 1, 2, 3, 4};
extern "C" {
const char* zb_last_error(void) { return zb::t_last_error.c_str(); }
int zb_index_options(zb_index* ix, zb_options* o) {
    if (ix != FAKE) return ZB_ERR_NO_DEVICE;
    memset(o, 0, sizeof *o);
    o->dim = 2;
    o->num_trees = 2;
    return ZB_OK;
}
int zb_index_load_forest(zb_index*, uint64_t, const float*, const uint8_t*, const int64_t*, const int32_t*, const int32_t*, const float*,
                         const float*, const int64_t*, const uint64_t*) { return ZB_ERR_NO_DEVICE; }
int zb_index_forest_sizes(zb_index* ix, int64_t* s) {
    if (ix != FAKE) return ZB_ERR_NO_DEVICE;
    s[0] = 6; s[1] = 2; s[2] = 4; s[3] = 10;
    return ZB_OK;
}
int zb_index_export_forest(zb_index* ix, int32_t* nodes, int32_t* roots, float* coef, float* cst, int64_t* leaf_off, uint64_t* members) {
    if (ix != FAKE) return ZB_ERR_NO_DEVICE;
    memcpy(nodes, F_NODES, sizeof F_NODES); memcpy(roots, F_ROOTS, sizeof F_ROOTS); memcpy(coef, F_COEF, sizeof F_COEF);
    memcpy(cst, F_CST, sizeof F_CST); memcpy(leaf_off, F_LEAF_OFF, sizeof F_LEAF_OFF); memcpy(members, F_MEMBERS, sizeof F_MEMBERS);
    return ZB_OK;
}
int zb_index_stats(zb_index* ix, zb_stats* st) {
    if (ix != FAKE) return ZB_ERR_NO_DEVICE;
    memset(st, 0, sizeof *st);
    st->total_rows = 5;
    return ZB_OK;
}
int zb_index_export_rows(zb_index* ix, uint64_t first, uint64_t n, float*, uint8_t* ids, uint8_t* live) {
    if (ix != FAKE || first != 0 || n != 5) return ZB_ERR_NO_DEVICE;
    for (uint64_t i = 0; i < n; ++i) {
        memset(ids + 16 * i, 0, 16);
        ids[16 * i + 15] = (uint8_t)(0xA0 + i);
        live[i] = i != 3;
    }
    return ZB_OK;
}
}

static int check_exports() {
    // expected: the same forest with row 3 left out, encoded tree by tree with the pure encoder
    const int64_t leaf_off[] = {0, 2, 3, 4, 8};
    uint8_t ids[8 * 16] = {0};
    const int who[] = {0, 1, 2, 4, 0, 1, 2, 4};
    for (int j = 0; j < 8; ++j) ids[16 * j + 15] = (uint8_t)(0xA0 + who[j]);
    std::vector<uint8_t> want[2], all;
    for (int t = 0; t < 2; ++t) {
        want[t].resize(1024);
        uint64_t need = 0;
        if (zb_tree_blob_encode(2, 6, F_NODES, F_ROOTS[t], F_COEF, F_CST, leaf_off, ids, want[t].data(), want[t].size(), &need) != ZB_OK) return 1;
        want[t].resize(need);
        all.insert(all.end(), want[t].begin(), want[t].end());
        uint64_t got_need = 0;
        if (zb_index_export_tree_blob(FAKE, (uint32_t)t, nullptr, 0, &got_need) != ZB_OK || got_need != need) return 2;
        std::vector<uint8_t> got(need);
        if (zb_index_export_tree_blob(FAKE, (uint32_t)t, got.data(), need, &got_need) != ZB_OK || got != want[t]) return 3;
        if (zb_index_export_tree_blob(FAKE, (uint32_t)t, got.data(), need - 1, &got_need) != ZB_ERR_INVALID) return 4;   // short buffer
    }
    uint64_t sizes[2] = {0, 0}, total = 0;
    if (zb_index_export_tree_blobs(FAKE, nullptr, 0, sizes, &total) != ZB_OK || total != all.size() || sizes[0] != want[0].size() ||
        sizes[1] != want[1].size()) return 5;
    std::vector<uint8_t> got(total);
    if (zb_index_export_tree_blobs(FAKE, got.data(), total, sizes, &total) != ZB_OK || got != all) return 6;
    if (zb_index_export_tree_blobs(FAKE, got.data(), total - 3, sizes, &total) != ZB_ERR_INVALID) return 7;
    uint64_t one = 0;
    if (zb_index_export_tree_blob(FAKE, 2, nullptr, 0, &one) != ZB_ERR_INVALID) return 8;                                 // no such tree
    return 0;
}

static uint64_t state = 0x9E3779B97F4A7C15ull;
static uint32_t rnd() {
    state ^= state << 13; state ^= state >> 7; state ^= state << 17;
    return (uint32_t)(state >> 16);
}

// a valid random tree blob of `dim`-dimensional planes over ids 1..nid (each id in exactly one leaf)
static void gen_tree(std::vector<uint8_t>& out, uint32_t dim, int depth, uint32_t& next_id, uint32_t nid) {
    auto put = [&](const void* p, size_t k) { out.insert(out.end(), (const uint8_t*)p, (const uint8_t*)p + k); };
    if (depth > 0 && rnd() % 3 != 0) {
        uint32_t tag = 0;
        put(&tag, 4);
        for (uint32_t i = 0; i <= dim; ++i) { float f = (float)(int)(rnd() % 2001 - 1000) / 100.0f; put(&f, 4); }
        gen_tree(out, dim, depth - 1, next_id, nid);
        gen_tree(out, dim, depth - 1, next_id, nid);
    } else {
        uint32_t tag = 1;
        uint64_t cnt = next_id <= nid ? rnd() % 4 : 0;
        if (next_id + cnt > nid + 1) cnt = nid + 1 - next_id;
        put(&tag, 4);
        put(&cnt, 8);
        for (uint64_t i = 0; i < cnt; ++i) {
            uint64_t len = 16;
            uint8_t id[16] = {0};
            uint32_t v = next_id++;
            id[12] = (uint8_t)(v >> 24); id[13] = (uint8_t)(v >> 16); id[14] = (uint8_t)(v >> 8); id[15] = (uint8_t)v;
            put(&len, 8);
            put(id, 16);
        }
    }
}

int main(int argc, char** argv) {
    const int iters = argc > 1 ? atoi(argv[1]) : 20000;
    if (int rc = check_exports()) {
        std::printf("export of the canned forest failed at check %d: %s\n", rc, zb_last_error());
        return 1;
    }
    long ok = 0, rejected = 0;
    for (int it = 0; it < iters; ++it) {
        const uint32_t dim = 1 + rnd() % 6, nid = rnd() % 24;
        uint32_t next_id = 1;
        std::vector<uint8_t> blob;
        gen_tree(blob, dim, 5, next_id, nid);
        std::vector<uint8_t> m = blob;
        const int kind = it % 5;
        if (kind == 1 && !m.empty()) m.resize(rnd() % m.size());                                   // truncate
        else if (kind == 2) for (int k = 0; k < 3 && !m.empty(); ++k) m[rnd() % m.size()] ^= (uint8_t)(1u << (rnd() % 8));  // bit flips
        else if (kind == 3) for (int k = 0; k < 8; ++k) m.push_back((uint8_t)rnd());              // trailing garbage
        else if (kind == 4 && m.size() > 12) { uint64_t huge = ~0ull >> (rnd() % 8); memcpy(&m[rnd() % (m.size() - 8)], &huge, 8); }  // absurd lengths
        int64_t s4[4] = {0, 0, 0, 0};
        int rc = zb_tree_blob_decode(dim, m.data(), m.size(), s4, nullptr, nullptr, nullptr, nullptr, nullptr);
        if (rc != ZB_OK && rc != ZB_ERR_INVALID) { std::printf("unexpected status %d\n", rc); return 1; }
        if (rc == ZB_OK) {
            std::vector<int32_t> nodes((size_t)s4[0] * 4 + 4);
            std::vector<float> coef((size_t)s4[1] * dim + 1), cst((size_t)s4[1] + 1);
            std::vector<int64_t> leaf_off((size_t)s4[2] + 1);
            std::vector<uint8_t> ids((size_t)s4[3] * 16 + 16);
            rc = zb_tree_blob_decode(dim, m.data(), m.size(), s4, nodes.data(), coef.data(), cst.data(), leaf_off.data(), ids.data());
            if (rc != ZB_OK) { std::printf("second pass disagrees with the first\n"); return 1; }
            std::vector<uint8_t> back(m.size() + 16);
            uint64_t need = 0;
            rc = zb_tree_blob_encode(dim, s4[0], nodes.data(), 0, coef.data(), cst.data(), leaf_off.data(), ids.data(), back.data(),
                                     back.size(), &need);
            if (rc != ZB_OK || need != m.size() || memcmp(back.data(), m.data(), m.size()) != 0) {
                std::printf("decode -> encode is not the identity on an accepted blob\n");
                return 1;
            }
            ++ok;
        } else {
            ++rejected;
        }
        // the whole-store path: the (possibly mutated) blob twice as a 2-tree store over ids 1..nid (some missing, some extra)
        std::vector<uint8_t> keys;
        const uint32_t have = nid ? rnd() % (nid + 2) : 0;
        for (uint32_t v = 1; v <= have; ++v) {
            uint8_t id[16] = {0};
            id[12] = (uint8_t)(v >> 24); id[13] = (uint8_t)(v >> 16); id[14] = (uint8_t)(v >> 8); id[15] = (uint8_t)v;
            keys.insert(keys.end(), id, id + 16);
        }
        const uint8_t* ptrs[2] = {m.data(), blob.data()};
        const uint64_t lens[2] = {m.size(), blob.size()};
        zb_flat_store* fs = nullptr;
        zb_import_report rep;
        rc = zb_store_flatten(dim, have, keys.data(), 2, ptrs, lens, &fs, &rep);
        if (rc != ZB_OK && rc != ZB_ERR_INVALID) { std::printf("flatten: unexpected status %d\n", rc); return 1; }
        if (rc == ZB_OK) {
            if (rep.rows_loaded + rep.orphan_rows != have) { std::printf("flatten: rows are neither loaded nor orphans\n"); return 1; }
            zb_flat_store_free(fs);
        }
        uint8_t z[44];
        for (auto& b : z) b = (uint8_t)rnd();
        zb_zebra_file_decode(z, 40 + (rnd() % 3) * 2, rnd() % 14, nullptr, nullptr, nullptr, nullptr);
    }
    std::printf("fuzz ok: %ld accepted, %ld rejected\n", ok, rejected);
    return 0;
}
