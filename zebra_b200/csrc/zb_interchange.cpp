// zb_interchange.cpp -- the reference's stored VALUES in and out of the device index.
//
// The reference (emmyoh/zebra) persists through fjall + bincode (/root/reference/src/database/index/lsh.rs:63-119):
//   embeddings partition   key = Uuid bytes, value = bincode(legacy) of Embedding<N>   (lsh.rs:91-97)
//   trees partition        key = Uuid bytes, value = bincode(legacy) of Node<N>        (lsh.rs:99-105, types :45-60)
//   <uuid>.zebra           bincode(legacy) of DatabaseInner                            (core.rs:19-29, :183-190)
// bincode 2 `config::legacy()` = little endian, fixed-width integers, u64 lengths, u32 enum variant index.  Through serde:
//   Embedding<N>   serde_with "[_; N]" -> a tuple of N f32, no length prefix (lib.rs:16-18)  = 4N bytes
//   Uuid           non-human-readable serializers get `serialize_bytes`                       = u64 16 | 16 bytes
//   Vec<Uuid>      u64 count | elements;   Box<T> is transparent;   unit structs are empty
//   Node::Inner    u32 0 | Hyperplane { coefficients 4N, constant f32 } | left_node | right_node   (lsh.rs:52-57)
//   Node::Leaf     u32 1 | Vec<Uuid>                                                               (lsh.rs:59-60)
// (bincode and uuid are un-vendored dependencies, Cargo.toml:56-58: the layout is restated from their published formats;
// oracle/zb_bincode.py restates it independently and tests/test_interchange.py pins hand-written byte strings.)
//
// Pure host code: the codecs need no device; import / export go through the public C ABI of this library
// (zb_index_load_forest / zb_index_export_forest / zb_index_export_rows), so they hold no private state.
#include <algorithm>
#include <cstring>
#include <memory>
#include <numeric>
#include <string>
#include <unordered_map>
#include <vector>

#include "zb_host.h"

namespace {

struct Reader {
    const uint8_t* p;
    uint64_t n, at = 0;
    void need(uint64_t k, const char* what) {
        if (k > n - at) throw zb::Error(ZB_ERR_INVALID, zb::fmt("tree blob truncated at byte %llu reading %s", (unsigned long long)at, what));
    }
    uint32_t u32(const char* what) {
        need(4, what);
        uint32_t v;
        memcpy(&v, p + at, 4);
        at += 4;
        return v;
    }
    uint64_t u64(const char* what) {
        need(8, what);
        uint64_t v;
        memcpy(&v, p + at, 8);
        at += 8;
        return v;
    }
    const uint8_t* raw(uint64_t k, const char* what) {
        need(k, what);
        const uint8_t* r = p + at;
        at += k;
        return r;
    }
};

struct Writer {
    uint8_t* out;  // may be NULL: size only
    uint64_t cap, at = 0;
    void put(const void* src, uint64_t k) {
        if (out && at + k <= cap) memcpy(out + at, src, k);
        at += k;
    }
    void u32(uint32_t v) { put(&v, 4); }
    void u64(uint64_t v) { put(&v, 8); }
};

struct FlatTree {
    std::vector<int32_t> nodes;  // 4 per node
    std::vector<float> coef, cst;
    std::vector<int64_t> leaf_off{0};
    std::vector<uint8_t> ids;  // 16 per member
    uint32_t max_depth = 0;
};

// Iterative preorder parse (a degenerate tree must not overflow the C stack).  `fill` = false: count only.
void parse_tree(uint32_t dim, const uint8_t* blob, uint64_t bytes, bool fill, FlatTree& ft, int64_t sizes4[4]) {
    Reader r{blob, bytes};
    struct Pending { int32_t parent; int side; uint32_t depth; };
    std::vector<Pending> stack{{-1, 0, 0}};
    int64_t n_nodes = 0, n_planes = 0, n_leaves = 0, n_members = 0;
    const uint64_t plane_bytes = (uint64_t)dim * 4;
    while (!stack.empty()) {
        const Pending pd = stack.back();
        stack.pop_back();
        const uint32_t tag = r.u32("the Node variant");
        if (n_nodes >= (int64_t)0x7FFFFFF0) throw zb::Error(ZB_ERR_INVALID, "tree blob has too many nodes");
        const int32_t id = (int32_t)n_nodes++;
        ft.max_depth = std::max(ft.max_depth, pd.depth);
        if (fill && pd.parent >= 0) ft.nodes[4 * (size_t)pd.parent + 1 + pd.side] = id;
        if (tag == 0) {
            const uint8_t* c = r.raw(plane_bytes, "hyperplane coefficients");
            const uint8_t* k = r.raw(4, "hyperplane constant");
            if (fill) {
                memcpy(ft.coef.data() + (size_t)n_planes * dim, c, plane_bytes);
                memcpy(ft.cst.data() + n_planes, k, 4);
                int32_t* nd = ft.nodes.data() + 4 * (size_t)id;
                nd[0] = (int32_t)n_planes; nd[1] = -1; nd[2] = -1; nd[3] = -1;
            }
            n_planes++;
            stack.push_back({id, 1, pd.depth + 1});  // right_node (above) is serialised after ...
            stack.push_back({id, 0, pd.depth + 1});  // ... left_node (below), so it is parsed second
        } else if (tag == 1) {
            const uint64_t cnt = r.u64("the leaf length");
            if (cnt > (bytes - r.at) / 24) throw zb::Error(ZB_ERR_INVALID, zb::fmt("leaf of %llu ids does not fit the blob", (unsigned long long)cnt));
            for (uint64_t i = 0; i < cnt; ++i) {
                const uint64_t len = r.u64("a Uuid length");
                if (len != 16) throw zb::Error(ZB_ERR_INVALID, zb::fmt("Uuid of %llu bytes at byte %llu", (unsigned long long)len, (unsigned long long)r.at));
                const uint8_t* idb = r.raw(16, "a Uuid");
                if (fill) memcpy(ft.ids.data() + 16 * (size_t)(n_members + (int64_t)i), idb, 16);
            }
            n_members += (int64_t)cnt;
            if (fill) {
                int32_t* nd = ft.nodes.data() + 4 * (size_t)id;
                nd[0] = -1; nd[1] = -1; nd[2] = -1; nd[3] = (int32_t)n_leaves;
                ft.leaf_off[(size_t)n_leaves + 1] = n_members;
            }
            n_leaves++;
        } else {
            throw zb::Error(ZB_ERR_INVALID, zb::fmt("Node variant %u at byte %llu (expected 0 = Inner or 1 = Leaf)", tag, (unsigned long long)(r.at - 4)));
        }
    }
    if (r.at != bytes) throw zb::Error(ZB_ERR_INVALID, zb::fmt("%llu trailing bytes after the tree", (unsigned long long)(bytes - r.at)));
    sizes4[0] = n_nodes; sizes4[1] = n_planes; sizes4[2] = n_leaves; sizes4[3] = n_members;
}

void decode_tree(uint32_t dim, const uint8_t* blob, uint64_t bytes, FlatTree& ft) {
    int64_t s[4];
    parse_tree(dim, blob, bytes, false, ft, s);
    ft.nodes.assign((size_t)s[0] * 4, -1);
    ft.coef.assign((size_t)s[1] * dim, 0.f);
    ft.cst.assign((size_t)s[1], 0.f);
    ft.leaf_off.assign((size_t)s[2] + 1, 0);
    ft.ids.assign((size_t)s[3] * 16, 0);
    parse_tree(dim, blob, bytes, true, ft, s);
}

// `keep`: optional per-member flag (0 = leave the id out of its leaf).
uint64_t encode_tree(uint32_t dim, int64_t n_nodes, const int32_t* nodes, int32_t root, const float* coef, const float* cst,
                     const int64_t* leaf_off, const uint8_t* ids16, const uint8_t* keep, uint8_t* out, uint64_t cap) {
    Writer w{out, cap};
    std::vector<int32_t> stack{root};
    int64_t visited = 0;
    while (!stack.empty()) {
        const int32_t i = stack.back();
        stack.pop_back();
        if (i < 0 || i >= n_nodes) throw zb::Error(ZB_ERR_INVALID, zb::fmt("node %d out of range", i));
        if (++visited > n_nodes) throw zb::Error(ZB_ERR_INVALID, "forest has a cycle");
        const int32_t* nd = nodes + 4 * (size_t)i;
        if (nd[0] >= 0) {
            w.u32(0);
            w.put(coef + (size_t)nd[0] * dim, (uint64_t)dim * 4);
            w.put(cst + nd[0], 4);
            stack.push_back(nd[2]);
            stack.push_back(nd[1]);  // left first
        } else {
            if (nd[3] < 0) throw zb::Error(ZB_ERR_INVALID, zb::fmt("node %d is neither inner nor leaf", i));
            const int64_t a = leaf_off[nd[3]], b = leaf_off[nd[3] + 1];
            uint64_t cnt = 0;
            for (int64_t j = a; j < b; ++j) cnt += (!keep || keep[j]) ? 1 : 0;
            w.u32(1);
            w.u64(cnt);
            for (int64_t j = a; j < b; ++j) {
                if (keep && !keep[j]) continue;
                w.u64(16);
                w.put(ids16 + 16 * (size_t)j, 16);
            }
        }
    }
    return w.at;
}

struct Key16 {
    uint64_t hi, lo;  // big-endian halves: integer order == byte order == Uuid's Ord
    bool operator==(const Key16& o) const { return hi == o.hi && lo == o.lo; }
    bool operator<(const Key16& o) const { return hi < o.hi || (hi == o.hi && lo < o.lo); }
};
struct Key16Hash {
    size_t operator()(const Key16& k) const { return (size_t)(k.hi * 0x9E3779B97F4A7C15ull ^ (k.lo + 0x7F4A7C15ull + (k.hi << 6))); }
};
Key16 load_key(const uint8_t* p) {
    Key16 k{0, 0};
    for (int i = 0; i < 8; ++i) k.hi = (k.hi << 8) | p[i];
    for (int i = 8; i < 16; ++i) k.lo = (k.lo << 8) | p[i];
    return k;
}

}  // namespace

// The flat form of a whole store: what zb_index_load_forest takes, plus the order the rows go in.
struct zb_flat_store {
    int64_t sizes4[4] = {0, 0, 0, 0};
    std::vector<int32_t> nodes, roots;
    std::vector<float> coef, cst;
    std::vector<int64_t> leaf_off;
    std::vector<uint64_t> members;      // ordinals (positions in row_order)
    std::vector<uint32_t> row_order;    // input row index of ordinal o: the rows kept, in id order
    std::vector<uint32_t> orphan_rows;  // input row indices not loaded (some tree does not hold them), in id order
    zb_import_report report{};
};

namespace {

bool metric_has_power(uint32_t metric) { return metric == ZB_METRIC_MINKOWSKI || metric == ZB_METRIC_PNORM; }

}  // namespace

extern "C" {

int zb_tree_blob_decode(uint32_t dim, const uint8_t* blob, uint64_t bytes, int64_t* sizes4, int32_t* nodes, float* coef,
                        float* cst, int64_t* leaf_off, uint8_t* member_ids16) {
    ZB_API_BEGIN
    ZB_REQUIRE(dim >= 1 && blob && sizes4, ZB_ERR_INVALID, "bad argument");
    FlatTree ft;
    if (!nodes) {
        parse_tree(dim, blob, bytes, false, ft, sizes4);
    } else {
        ZB_REQUIRE(leaf_off && (coef && cst), ZB_ERR_INVALID, "output arrays missing");
        decode_tree(dim, blob, bytes, ft);
        sizes4[0] = (int64_t)ft.nodes.size() / 4; sizes4[1] = (int64_t)ft.cst.size();
        sizes4[2] = (int64_t)ft.leaf_off.size() - 1; sizes4[3] = (int64_t)ft.ids.size() / 16;
        memcpy(nodes, ft.nodes.data(), ft.nodes.size() * 4);
        if (!ft.cst.empty()) {
            memcpy(coef, ft.coef.data(), ft.coef.size() * 4);
            memcpy(cst, ft.cst.data(), ft.cst.size() * 4);
        }
        memcpy(leaf_off, ft.leaf_off.data(), ft.leaf_off.size() * 8);
        if (!ft.ids.empty()) {
            ZB_REQUIRE(member_ids16, ZB_ERR_INVALID, "member_ids16 is NULL");
            memcpy(member_ids16, ft.ids.data(), ft.ids.size());
        }
    }
    ZB_API_END
}

int zb_tree_blob_encode(uint32_t dim, int64_t n_nodes, const int32_t* nodes, int32_t root, const float* coef,
                        const float* cst, const int64_t* leaf_off, const uint8_t* member_ids16, uint8_t* out,
                        uint64_t cap, uint64_t* out_bytes) {
    ZB_API_BEGIN
    ZB_REQUIRE(dim >= 1 && nodes && leaf_off && out_bytes && n_nodes >= 1, ZB_ERR_INVALID, "bad argument");
    const uint64_t need = encode_tree(dim, n_nodes, nodes, root, coef, cst, leaf_off, member_ids16, nullptr, out, cap);
    *out_bytes = need;
    ZB_REQUIRE(!out || need <= cap, ZB_ERR_INVALID, "tree blob needs %llu bytes, buffer holds %llu", (unsigned long long)need,
               (unsigned long long)cap);
    ZB_API_END
}

int zb_zebra_file_encode(const uint8_t* uuid16, uint32_t metric, int32_t power, uint64_t max_node_size,
                         uint64_t num_trees, uint8_t* out, uint64_t cap, uint64_t* out_bytes) {
    ZB_API_BEGIN
    ZB_REQUIRE(uuid16 && out_bytes && metric < ZB_METRIC_COUNT, ZB_ERR_INVALID, "bad argument");
    Writer w{out, cap};
    w.u64(16);
    w.put(uuid16, 16);                          // uuid: Uuid                      core.rs:25
    /* model: a unit struct, 0 bytes (model/text.rs:11, image.rs:50)                 core.rs:26 */
    if (metric_has_power(metric)) w.put(&power, 4);  // metric: unit struct or { power: i32 }   core.rs:27, distance.rs:162-165
    w.u64(max_node_size);                       // index_options.max_node_size: usize  lsh.rs:124-129
    w.u64(num_trees);
    *out_bytes = w.at;
    ZB_REQUIRE(!out || w.at <= cap, ZB_ERR_INVALID, ".zebra file needs %llu bytes", (unsigned long long)w.at);
    ZB_API_END
}

int zb_zebra_file_decode(const uint8_t* data, uint64_t bytes, uint32_t metric, uint8_t* out_uuid16, int32_t* out_power,
                         uint64_t* out_max_node_size, uint64_t* out_num_trees) {
    ZB_API_BEGIN
    ZB_REQUIRE(data && metric < ZB_METRIC_COUNT, ZB_ERR_INVALID, "bad argument");
    const uint64_t want = 24 + (metric_has_power(metric) ? 4 : 0) + 16;
    ZB_REQUIRE(bytes == want, ZB_ERR_INVALID, ".zebra file of %llu bytes, expected %llu for this metric", (unsigned long long)bytes,
               (unsigned long long)want);
    Reader r{data, bytes};
    ZB_REQUIRE(r.u64("the uuid length") == 16, ZB_ERR_INVALID, ".zebra file: uuid length is not 16");
    const uint8_t* id = r.raw(16, "the uuid");
    if (out_uuid16) memcpy(out_uuid16, id, 16);
    int32_t power = 0;
    if (metric_has_power(metric)) memcpy(&power, r.raw(4, "the metric power"), 4);
    if (out_power) *out_power = power;
    const uint64_t mns = r.u64("max_node_size"), nt = r.u64("num_trees");
    if (out_max_node_size) *out_max_node_size = mns;
    if (out_num_trees) *out_num_trees = nt;
    ZB_API_END
}

int zb_store_flatten(uint32_t dim, uint64_t n, const uint8_t* ids16, uint32_t n_trees, const uint8_t* const* tree_blobs,
                     const uint64_t* tree_blob_bytes, zb_flat_store** out, zb_import_report* report) {
    ZB_API_BEGIN
    ZB_REQUIRE(dim >= 1 && out && (ids16 || !n) && tree_blobs && tree_blob_bytes && n_trees >= 1, ZB_ERR_INVALID, "bad argument");
    ZB_REQUIRE(n < (1ull << 32), ZB_ERR_INVALID, "too many rows for one import");
    std::unique_ptr<zb_flat_store> fs(new zb_flat_store());
    // ---- rows in id order (the key order of the reference's store): ordinal = rank of the id ----
    std::vector<Key16> keys(n);
    for (uint64_t i = 0; i < n; ++i) keys[i] = load_key(ids16 + 16 * i);
    std::vector<uint32_t> perm(n);
    std::iota(perm.begin(), perm.end(), 0u);
    if (!std::is_sorted(keys.begin(), keys.end()))
        std::sort(perm.begin(), perm.end(), [&](uint32_t a, uint32_t b) { return keys[a] < keys[b]; });
    std::unordered_map<Key16, uint32_t, Key16Hash> rank_of;  // id -> position in id order
    rank_of.reserve((size_t)n * 2);
    for (uint64_t r = 0; r < n; ++r) {
        const bool fresh = rank_of.emplace(keys[perm[r]], (uint32_t)r).second;
        ZB_REQUIRE(fresh, ZB_ERR_INVALID, "id of row %u appears twice in the store", perm[r]);
    }
    // ---- trees: decode, map ids to ranks, check that every row sits in exactly one leaf of every tree ----
    std::vector<FlatTree> trees(n_trees);
    std::vector<std::vector<uint32_t>> member_rank(n_trees);  // per member: rank, or 0xFFFFFFFF = id without embedding
    std::vector<uint8_t> seen((size_t)n);
    std::vector<uint8_t> orphan((size_t)n, 0);
    uint64_t missing = 0;
    uint32_t max_depth = 0;
    for (uint32_t t = 0; t < n_trees; ++t) {
        ZB_REQUIRE(tree_blobs[t], ZB_ERR_INVALID, "tree blob %u is NULL", t);
        decode_tree(dim, tree_blobs[t], tree_blob_bytes[t], trees[t]);
        max_depth = std::max(max_depth, trees[t].max_depth);
        const size_t nm = trees[t].ids.size() / 16;
        member_rank[t].resize(nm);
        std::fill(seen.begin(), seen.end(), 0);
        for (size_t j = 0; j < nm; ++j) {
            auto it = rank_of.find(load_key(trees[t].ids.data() + 16 * j));
            if (it == rank_of.end()) {
                member_rank[t][j] = 0xFFFFFFFFu;
                missing++;
                continue;
            }
            ZB_REQUIRE(!seen[it->second], ZB_ERR_INVALID, "tree %u holds the id of row %u twice", t, perm[it->second]);
            seen[it->second] = 1;
            member_rank[t][j] = it->second;
        }
        for (uint64_t r = 0; r < n; ++r)
            if (!seen[r]) orphan[r] = 1;
    }
    // ---- drop orphans: final ordinal = rank among the rows kept ----
    std::vector<uint32_t> ordinal_of_rank((size_t)n, 0xFFFFFFFFu);
    for (uint64_t r = 0; r < n; ++r) {
        if (orphan[r]) {
            fs->orphan_rows.push_back(perm[r]);
        } else {
            ordinal_of_rank[r] = (uint32_t)fs->row_order.size();
            fs->row_order.push_back(perm[r]);
        }
    }
    // ---- one flat forest (tree after tree: preorder inside a tree, as zb_index_export_forest numbers it) ----
    int64_t* tot = fs->sizes4;
    for (auto& ft : trees) {
        tot[0] += (int64_t)ft.nodes.size() / 4;
        tot[1] += (int64_t)ft.cst.size();
        tot[2] += (int64_t)ft.leaf_off.size() - 1;
    }
    ZB_REQUIRE(tot[0] < 0x7FFFFFF0ll, ZB_ERR_INVALID, "the forest has too many nodes");
    fs->nodes.resize((size_t)tot[0] * 4);
    fs->roots.resize(n_trees);
    fs->coef.resize((size_t)tot[1] * dim + 1);   // (+1: never an empty buffer behind the pointer)
    fs->cst.resize((size_t)tot[1] + 1);
    fs->leaf_off.assign((size_t)tot[2] + 1, 0);
    int64_t nb = 0, pb = 0, lb = 0;
    for (uint32_t t = 0; t < n_trees; ++t) {
        const FlatTree& ft = trees[t];
        const int64_t nn = (int64_t)ft.nodes.size() / 4, np = (int64_t)ft.cst.size(), nl = (int64_t)ft.leaf_off.size() - 1;
        fs->roots[t] = (int32_t)nb;
        for (int64_t i = 0; i < nn; ++i) {
            const int32_t* s = ft.nodes.data() + 4 * i;
            int32_t* d = fs->nodes.data() + 4 * (nb + i);
            if (s[0] >= 0) {
                d[0] = (int32_t)(pb + s[0]); d[1] = (int32_t)(nb + s[1]); d[2] = (int32_t)(nb + s[2]); d[3] = -1;
            } else {
                d[0] = -1; d[1] = -1; d[2] = -1; d[3] = (int32_t)(lb + s[3]);
            }
        }
        if (np) {
            memcpy(fs->coef.data() + (size_t)pb * dim, ft.coef.data(), (size_t)np * dim * 4);
            memcpy(fs->cst.data() + pb, ft.cst.data(), (size_t)np * 4);
        }
        for (int64_t l = 0; l < nl; ++l) {
            for (int64_t j = ft.leaf_off[l]; j < ft.leaf_off[l + 1]; ++j) {
                const uint32_t r = member_rank[t][j];
                if (r == 0xFFFFFFFFu || orphan[r]) continue;
                fs->members.push_back(ordinal_of_rank[r]);
            }
            fs->leaf_off[(size_t)(lb + l) + 1] = (int64_t)fs->members.size();
        }
        nb += nn; pb += np; lb += nl;
    }
    tot[3] = (int64_t)fs->members.size();
    fs->members.push_back(0);  // never an empty buffer behind the pointer
    fs->report.rows_loaded = fs->row_order.size();
    fs->report.missing_ids = missing;
    fs->report.orphan_rows = fs->orphan_rows.size();
    fs->report.nodes = (uint64_t)tot[0];
    fs->report.planes = (uint64_t)tot[1];
    fs->report.leaves = (uint64_t)tot[2];
    fs->report.max_depth = max_depth;
    if (report) *report = fs->report;
    *out = fs.release();
    ZB_API_END
}

int zb_flat_store_view(const zb_flat_store* fs, int64_t* sizes4, const int32_t** nodes, const int32_t** roots, const float** coef,
                       const float** cst, const int64_t** leaf_off, const uint64_t** members, const uint32_t** row_order,
                       const uint32_t** orphan_rows) {
    ZB_API_BEGIN
    ZB_REQUIRE(fs, ZB_ERR_INVALID, "NULL argument");
    if (sizes4) memcpy(sizes4, fs->sizes4, sizeof fs->sizes4);
    if (nodes) *nodes = fs->nodes.data();
    if (roots) *roots = fs->roots.data();
    if (coef) *coef = fs->coef.data();
    if (cst) *cst = fs->cst.data();
    if (leaf_off) *leaf_off = fs->leaf_off.data();
    if (members) *members = fs->members.data();
    if (row_order) *row_order = fs->row_order.data();
    if (orphan_rows) *orphan_rows = fs->orphan_rows.data();
    ZB_API_END
}

int zb_flat_store_free(zb_flat_store* fs) {
    ZB_API_BEGIN
    delete fs;
    ZB_API_END
}

int zb_index_import_store(zb_index* index, uint64_t n, const uint8_t* ids16, const float* rows, uint32_t n_trees,
                          const uint8_t* const* tree_blobs, const uint64_t* tree_blob_bytes, zb_import_report* report,
                          uint8_t* out_orphan_ids16, uint64_t orphan_cap) {
    ZB_API_BEGIN
    ZB_REQUIRE(index && (rows || !n), ZB_ERR_INVALID, "bad argument");
    zb_options opt;
    int rc = zb_index_options(index, &opt);
    if (rc != ZB_OK) return rc;
    const uint32_t dim = opt.dim;
    ZB_REQUIRE(n_trees == opt.num_trees, ZB_ERR_INVALID, "the store holds %u trees, the index was created for %u", n_trees, opt.num_trees);
    zb_flat_store* raw = nullptr;
    rc = zb_store_flatten(dim, n, ids16, n_trees, tree_blobs, tree_blob_bytes, &raw, nullptr);
    if (rc != ZB_OK) return rc;  // message already set
    std::unique_ptr<zb_flat_store> fs(raw);
    const uint64_t n_keep = fs->row_order.size();
    bool identity = n_keep == n;
    for (uint64_t i = 0; identity && i < n_keep; ++i) identity = fs->row_order[i] == i;
    std::vector<float> rows_kept;
    std::vector<uint8_t> ids_kept;
    if (!identity) {  // gather the kept rows into ordinal (= id) order
        rows_kept.resize((size_t)n_keep * dim + 1);
        ids_kept.resize((size_t)n_keep * 16 + 1);
        for (uint64_t o = 0; o < n_keep; ++o) {
            memcpy(rows_kept.data() + (size_t)o * dim, rows + (size_t)fs->row_order[o] * dim, (size_t)dim * 4);
            memcpy(ids_kept.data() + 16 * (size_t)o, ids16 + 16 * (size_t)fs->row_order[o], 16);
        }
    }
    for (uint64_t i = 0; out_orphan_ids16 && i < fs->orphan_rows.size() && i < orphan_cap; ++i)
        memcpy(out_orphan_ids16 + 16 * i, ids16 + 16 * (size_t)fs->orphan_rows[i], 16);
    rc = zb_index_load_forest(index, n_keep, identity ? rows : rows_kept.data(), identity ? ids16 : ids_kept.data(), fs->sizes4,
                              fs->nodes.data(), fs->roots.data(), fs->coef.data(), fs->cst.data(), fs->leaf_off.data(),
                              fs->members.data());
    if (rc != ZB_OK) return rc;  // message already set by the callee
    if (report) *report = fs->report;
    ZB_API_END
}

}  // extern "C"

namespace {
// The forest as flat arrays with member ids and live flags: what both export entry points encode from.
struct Exported {
    zb_options opt;
    int64_t s[4] = {0, 0, 0, 0};
    std::vector<int32_t> nodes, roots;
    std::vector<float> coef, cst;
    std::vector<int64_t> leaf_off;
    std::vector<uint8_t> mid, keep;
};
int export_forest_with_ids(zb_index* index, Exported& e) {
    int rc = zb_index_options(index, &e.opt);
    if (rc != ZB_OK) return rc;
    rc = zb_index_forest_sizes(index, e.s);
    if (rc != ZB_OK) return rc;
    ZB_REQUIRE(e.s[0] >= 1, ZB_ERR_STATE, "the index has no trees yet");
    e.nodes.resize((size_t)e.s[0] * 4);
    e.roots.resize(e.opt.num_trees);
    e.coef.resize((size_t)std::max<int64_t>(1, e.s[1]) * e.opt.dim);
    e.cst.resize((size_t)std::max<int64_t>(1, e.s[1]));
    e.leaf_off.resize((size_t)e.s[2] + 1);
    std::vector<uint64_t> members((size_t)std::max<int64_t>(1, e.s[3]));
    rc = zb_index_export_forest(index, e.nodes.data(), e.roots.data(), e.coef.data(), e.cst.data(), e.leaf_off.data(), members.data());
    if (rc != ZB_OK) return rc;
    zb_stats st;
    rc = zb_index_stats(index, &st);
    if (rc != ZB_OK) return rc;
    // ids and live flags of every row (removed rows leave the leaves: DESIGN.md D1)
    std::vector<uint8_t> ids((size_t)std::max<uint64_t>(1, st.total_rows) * 16), live((size_t)std::max<uint64_t>(1, st.total_rows));
    rc = zb_index_export_rows(index, 0, st.total_rows, nullptr, ids.data(), live.data());
    if (rc != ZB_OK) return rc;
    e.mid.resize((size_t)std::max<int64_t>(1, e.s[3]) * 16);
    e.keep.resize((size_t)std::max<int64_t>(1, e.s[3]));
    for (int64_t j = 0; j < e.s[3]; ++j) {
        ZB_REQUIRE(members[j] < st.total_rows, ZB_ERR_STATE, "member ordinal out of range");
        memcpy(e.mid.data() + 16 * (size_t)j, ids.data() + 16 * (size_t)members[j], 16);
        e.keep[j] = live[members[j]];
    }
    return ZB_OK;
}
}  // namespace

extern "C" {

int zb_index_export_tree_blob(zb_index* index, uint32_t tree, uint8_t* out, uint64_t cap, uint64_t* out_bytes) {
    ZB_API_BEGIN
    ZB_REQUIRE(index && out_bytes, ZB_ERR_INVALID, "NULL argument");
    Exported e;
    int rc = export_forest_with_ids(index, e);
    if (rc != ZB_OK) return rc;
    ZB_REQUIRE(tree < e.opt.num_trees, ZB_ERR_INVALID, "tree %u out of range (%u trees)", tree, e.opt.num_trees);
    const uint64_t need = encode_tree(e.opt.dim, e.s[0], e.nodes.data(), e.roots[tree], e.coef.data(), e.cst.data(), e.leaf_off.data(),
                                      e.mid.data(), e.keep.data(), out, cap);
    *out_bytes = need;
    ZB_REQUIRE(!out || need <= cap, ZB_ERR_INVALID, "tree blob needs %llu bytes, buffer holds %llu", (unsigned long long)need,
               (unsigned long long)cap);
    ZB_API_END
}

int zb_index_export_tree_blobs(zb_index* index, uint8_t* out, uint64_t cap, uint64_t* out_blob_bytes, uint64_t* out_total) {
    ZB_API_BEGIN
    ZB_REQUIRE(index && out_total, ZB_ERR_INVALID, "NULL argument");
    Exported e;
    int rc = export_forest_with_ids(index, e);  // ONE export of the forest for all trees
    if (rc != ZB_OK) return rc;
    uint64_t at = 0;
    for (uint32_t t = 0; t < e.opt.num_trees; ++t) {
        const uint64_t room = out && at < cap ? cap - at : 0;
        const uint64_t need = encode_tree(e.opt.dim, e.s[0], e.nodes.data(), e.roots[t], e.coef.data(), e.cst.data(), e.leaf_off.data(),
                                          e.mid.data(), e.keep.data(), room ? out + at : nullptr, room);
        if (out_blob_bytes) out_blob_bytes[t] = need;
        at += need;
    }
    *out_total = at;
    ZB_REQUIRE(!out || at <= cap, ZB_ERR_INVALID, "the tree blobs need %llu bytes, buffer holds %llu", (unsigned long long)at,
               (unsigned long long)cap);
    ZB_API_END
}

}  // extern "C"
