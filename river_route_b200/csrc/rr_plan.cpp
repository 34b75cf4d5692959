// Host-side (CPU, no CUDA) part of librr_b200: topology validation, basin labelling, the
// wavefront plan (upstream-CSR, 32-reach blocks, systolic delays, block dependency DAG, ticket
// schedule) and the synthetic network generator.  Reference citations are in include/rr_b200.h.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <numeric>

#include "rr_internal.h"

static thread_local std::string g_err;
void rr_set_error(const std::string &msg) { g_err = msg; }
extern "C" const char *rr_last_error(void) { return g_err.c_str(); }
extern "C" int rr_version(void) { return 100; }

// ------------------------------------------------------------------------------------------
// id -> index hash (open addressing, splitmix finaliser)
// ------------------------------------------------------------------------------------------
namespace {
inline uint64_t mix64(uint64_t x) {
    x += 0x9e3779b97f4a7c15ull;
    x = (x ^ (x >> 30)) * 0xbf58476d1ce4e5b9ull;
    x = (x ^ (x >> 27)) * 0x94d049bb133111ebull;
    return x ^ (x >> 31);
}
struct IdTable {
    std::vector<int64_t> key;
    std::vector<int32_t> val;
    uint64_t mask = 0;
    explicit IdTable(int64_t n) {
        uint64_t cap = 16;
        while (cap < (uint64_t)n * 2 + 2) cap <<= 1;
        key.assign(cap, 0);
        val.assign(cap, -1);
        mask = cap - 1;
    }
    // returns previous index if the id was already present, else -1
    int32_t insert(int64_t id, int32_t idx) {
        uint64_t h = mix64((uint64_t)id) & mask;
        while (val[h] >= 0) {
            if (key[h] == id) return val[h];
            h = (h + 1) & mask;
        }
        key[h] = id;
        val[h] = idx;
        return -1;
    }
    int32_t find(int64_t id) const {
        uint64_t h = mix64((uint64_t)id) & mask;
        while (val[h] >= 0) {
            if (key[h] == id) return val[h];
            h = (h + 1) & mask;
        }
        return -1;
    }
};
}  // namespace

extern "C" int rr_downstream_index(int64_t n, const int64_t *river_ids, const int64_t *downstream_ids,
                                   int32_t *down_idx, int64_t *bad) {
    if (n < 0 || n > 0x7fffffff) { rr_set_error("reach count out of int32 range"); return 100; }
    IdTable tab(n);
    for (int64_t i = 0; i < n; ++i) {
        if (tab.insert(river_ids[i], (int32_t)i) >= 0) {
            if (bad) *bad = river_ids[i];
            rr_set_error("params_file contains duplicate river IDs.");
            return 1;
        }
    }
    for (int64_t i = 0; i < n; ++i) {
        const int64_t d = downstream_ids[i];
        if (d < 0) { down_idx[i] = -1; continue; }
        const int32_t di = tab.find(d);
        if (di < 0) {
            if (bad) *bad = d;
            rr_set_error("Unknown downstream_river_id: " + std::to_string(d));
            return 2;
        }
        if (di <= (int32_t)i) {
            if (bad) *bad = i;
            rr_set_error("params_file must be topologically sorted upstream to downstream");
            return 3;
        }
        down_idx[i] = di;
    }
    return 0;
}

// ------------------------------------------------------------------------------------------
// Plan
// ------------------------------------------------------------------------------------------
// Builds every derived structure of the plan from p->down (the working order; p->inv maps user ->
// working index when the plan is renumbered).
static int build_structures(rr_plan *p) {
    const int64_t n = p->n_work;   // working slots (padding slots are isolated reaches: no upstream, no downstream)
    const int32_t *down = p->down.data();
    const int64_t nb = p->n_blocks;
    p->n_edges = p->n_export = p->n_internal = p->n_outlets = 0;
    p->max_level = p->max_skew = p->max_deg = 0;

    // upstream-CSR: counting sort by downstream index keeps upstream indices ascending per row
    p->up_ptr.assign(n + 1, 0);
    for (int64_t i = 0; i < n; ++i)
        if (down[i] >= 0) { p->up_ptr[down[i] + 1]++; p->n_edges++; } else if (p->perm.empty() || p->perm[i] >= 0) p->n_outlets++;
    for (int64_t i = 0; i < n; ++i) p->up_ptr[i + 1] += p->up_ptr[i];
    p->up_idx.resize(p->n_edges);
    {
        // rows are filled in ascending USER index so that every confluence sums its inflows in the
        // reference's order even when the plan works on renumbered reaches
        std::vector<int32_t> fill(p->up_ptr.begin(), p->up_ptr.end() - 1);
        for (int64_t u = 0; u < p->n; ++u) {
            const int64_t i = p->inv.empty() ? u : p->inv[u];
            if (down[i] >= 0) p->up_idx[fill[down[i]]++] = (int32_t)i;
        }
    }
    p->is_hw.resize(n);
    p->n_hw = 0;
    for (int64_t i = 0; i < n; ++i) {
        p->is_hw[i] = p->up_ptr[i + 1] == p->up_ptr[i];
        if (p->is_hw[i] && (p->perm.empty() || p->perm[i] >= 0)) p->n_hw++;
    }

    // in-block systolic delays: every in-block edge gets lag exactly 1
    //   height[i] = longest in-block chain ending at i; roots (reaches whose downstream is in
    //   another block) start at skew = height, every in-block upstream is one less than its
    //   downstream, so a consumer lane always finds the value produced one step earlier.
    std::vector<uint8_t> height(n, 0);
    auto blk = [](int64_t i) { return i / RR_BLOCK; };
    for (int64_t i = 0; i < n; ++i) {
        const int32_t d = down[i];
        if (d >= 0 && blk(d) == blk(i)) height[d] = std::max<uint8_t>(height[d], height[i] + 1);
    }
    p->skew.resize(n);
    for (int64_t i = n - 1; i >= 0; --i) {
        const int32_t d = down[i];
        if (d >= 0 && blk(d) == blk(i)) { p->skew[i] = p->skew[d] - 1; p->n_internal++; }
        else p->skew[i] = height[i];
    }

    // exchange-buffer ids: only reaches whose downstream sits in another block export a series
    p->export_id.assign(n, -1);
    for (int64_t i = 0; i < n; ++i) {
        const int32_t d = down[i];
        if (d >= 0 && blk(d) != blk(i)) p->export_id[i] = (int32_t)p->n_export++;
    }

    // encoded upstream slots + block metadata + block dependency lists
    p->slot_src.resize(p->n_edges);
    p->meta.assign(nb, rr_blk_meta{0, 0, 0, 0});
    p->dep_ptr.assign(nb + 1, 0);
    p->dep_idx.clear();
    std::vector<int32_t> tmp;
    for (int64_t b = 0; b < nb; ++b) {
        rr_blk_meta &m = p->meta[b];
        tmp.clear();
        const int64_t lo = b * RR_BLOCK, hi = std::min<int64_t>(n, lo + RR_BLOCK);
        for (int64_t i = lo; i < hi; ++i) {
            const int32_t deg = p->up_ptr[i + 1] - p->up_ptr[i];
            if (deg > 65535) { rr_set_error("in-degree above 65535 is not supported"); return 100; }
            m.max_skew = std::max(m.max_skew, p->skew[i]);
            m.max_deg = std::max<uint16_t>(m.max_deg, (uint16_t)deg);
            for (int32_t k = 0; k < deg; ++k) {
                const int32_t e = p->up_ptr[i] + k;
                const int32_t u = p->up_idx[e];
                if (blk(u) == b) {
                    p->slot_src[e] = -(int32_t)(u - lo + 1) - (p->is_hw[u] ? 64 : 0);
                    m.int_mask |= (uint8_t)(k < RR_MAX_FAST_DEG ? (1u << k) : 0x80u);
                } else {
                    p->slot_src[e] = p->export_id[u] | (p->is_hw[u] ? RR_SLOT_HW_BIT : 0);
                    tmp.push_back((int32_t)blk(u));
                }
            }
        }
        std::sort(tmp.begin(), tmp.end());
        tmp.erase(std::unique(tmp.begin(), tmp.end()), tmp.end());
        int32_t lvl = 0;
        for (int32_t ub : tmp) lvl = std::max(lvl, p->meta[ub].level + 1);
        m.level = lvl;
        if (m.int_mask == 0 && m.max_skew == 0 && m.max_deg <= RR_MAX_FAST_DEG) m.int_mask |= RR_META_FAST;  // fast-path eligible
        p->dep_idx.insert(p->dep_idx.end(), tmp.begin(), tmp.end());
        p->dep_ptr[b + 1] = (int32_t)p->dep_idx.size();
        p->max_level = std::max(p->max_level, lvl);
        p->max_skew = std::max<int32_t>(p->max_skew, m.max_skew);
        p->max_deg = std::max<int32_t>(p->max_deg, m.max_deg);
    }
    p->blk_level.resize(nb);
    for (int64_t b = 0; b < nb; ++b) p->blk_level[b] = p->meta[b].level;

    // span of every exported series in block levels (producer block -> its single consumer block);
    // it sizes that series' ring in the exchange buffer (rr_build_schedule)
    p->exp_span.resize(p->n_export);
    for (int64_t i = 0; i < n; ++i)
        if (p->export_id[i] >= 0) p->exp_span[p->export_id[i]] = p->blk_level[blk(down[i])] - p->blk_level[blk(i)];

    // level buckets
    p->lvl_ptr.assign(p->max_level + 2, 0);
    for (int64_t b = 0; b < nb; ++b) p->lvl_ptr[p->blk_level[b] + 1]++;
    for (int32_t l = 0; l <= p->max_level; ++l) p->lvl_ptr[l + 1] += p->lvl_ptr[l];
    p->lvl_blk.resize(nb);
    {
        std::vector<int32_t> fill(p->lvl_ptr.begin(), p->lvl_ptr.end() - 1);
        for (int64_t b = 0; b < nb; ++b) p->lvl_blk[fill[p->blk_level[b]]++] = (int32_t)b;
    }
    p->all_fast = true;
    int64_t narrow = RR_NARROW_BLOCKS;
    if (const char *env = getenv("RR_NARROW_BLOCKS")) narrow = atoll(env);   // measurements only
    for (int64_t b = 0; b < nb; ++b) {
        rr_blk_meta &m = p->meta[b];
        if (!(m.int_mask & RR_META_FAST)) p->all_fast = false;
        if (p->lvl_ptr[m.level + 1] - p->lvl_ptr[m.level] < narrow) m.int_mask |= RR_META_NARROW;
    }
    return 0;
}

extern "C" int rr_plan_create(int64_t n, const int32_t *down, const rr_plan_opts *opts, rr_plan **out) {
    if (!out) { rr_set_error("null output pointer"); return 100; }
    *out = nullptr;
    if (n <= 0 || n > 0x7ffffff0ll) { rr_set_error("reach count must be in [1, 2^31)"); return 100; }
    for (int64_t i = 0; i < n; ++i) {
        if (down[i] >= 0 && (down[i] <= i || down[i] >= n)) {
            rr_set_error("params_file must be topologically sorted upstream to downstream");
            return 3;
        }
    }
    rr_plan *p = new rr_plan();
    if (opts) p->opts = *opts;
    p->auto_tile = p->opts.time_tile <= 0;
    if (p->auto_tile) p->opts.time_tile = 64;
    if (p->opts.threads_per_cta <= 0) p->opts.threads_per_cta = 256;
    p->opts.threads_per_cta = std::max(32, (p->opts.threads_per_cta / 32) * 32);
    if (p->opts.raw_budget_bytes <= 0) p->opts.raw_budget_bytes = 16ll << 30;
    if (!opts || opts->device < 0) p->opts.device = -1;
    p->n = n;
    p->n_work = n;
    p->n_blocks = (n + RR_BLOCK - 1) / RR_BLOCK;
    p->down.assign(down, down + n);
    // topological level of every reach (0 = headwater) in the user's order
    std::vector<int32_t> lvl(n, 0);
    int32_t depth = 0;
    for (int64_t i = 0; i < n; ++i) {
        if (down[i] >= 0) lvl[down[i]] = std::max(lvl[down[i]], lvl[i] + 1);
        depth = std::max(depth, lvl[i] + 1);
    }
    p->reach_depth = depth;
    int rc = 0;
    bool renumber = p->opts.renumber == 2;
    if (p->opts.renumber != 2) {
        rc = build_structures(p);
        if (rc) { delete p; return rc; }
        // Blocks of 32 consecutive reaches merge their dependencies; in an arbitrary (valid) order the
        // block DAG can be tens of times deeper than the river network.  Sorting reaches by level
        // makes it as shallow as the network itself.
        // Two reasons to work on level-sorted reaches instead (the caller's arrays stay as they are):
        //  * the block DAG is much deeper than the network, so launches are dependency-latency bound;
        //  * many blocks have in-block edges and miss the register-blocked fast path of the kernel.
        int64_t fast_blocks = 0;
        for (const rr_blk_meta &m : p->meta) fast_blocks += (m.int_mask & 0x40) ? 1 : 0;
        if (p->opts.renumber == 0 &&
            (p->max_level > 2 * (int64_t)depth + 16 || fast_blocks * 10 < p->n_blocks * 9)) renumber = true;
    }
    if (renumber) {
        // working order = stable sort by level, every level padded to whole 32-slot blocks: a block then holds reaches
        // of ONE level, which never connect, so no block has in-block edges and all of them run the register-blocked
        // fast path (deep levels with a handful of reaches each used to share blocks, chained inside them, and fell
        // back to the systolic path).  Padding slots are isolated dummy reaches (perm = -1, coefficients 0); at most
        // 31 per level.  perm[k] = user index of working slot k
        std::vector<int64_t> count(depth + 1, 0), start(depth + 1, 0);
        for (int64_t i = 0; i < n; ++i) count[lvl[i]]++;
        int64_t off = 0;
        for (int32_t l = 0; l < depth; ++l) { start[l] = off; off += (count[l] + RR_BLOCK - 1) / RR_BLOCK * RR_BLOCK; }
        if (off > 0x7ffffff0ll) { rr_set_error("reach count must be in [1, 2^31)"); delete p; return 100; }
        p->n_work = off;
        p->lvl0_slots = depth > 1 ? start[1] : off;
        p->n_blocks = off / RR_BLOCK;
        p->perm.assign(off, -1);
        p->inv.resize(n);
        for (int64_t i = 0; i < n; ++i) {
            const int64_t k = start[lvl[i]]++;
            p->perm[k] = (int32_t)i;
            p->inv[i] = (int32_t)k;
        }
        p->down.assign(off, -1);
        for (int64_t k = 0; k < off; ++k) {
            if (p->perm[k] < 0) continue;
            const int32_t d = down[p->perm[k]];
            p->down[k] = d >= 0 ? p->inv[d] : -1;
        }
        rc = build_structures(p);
        if (rc) { delete p; return rc; }
    }
    *out = p;
    return 0;
}

extern "C" void rr_plan_destroy(rr_plan *p) {
    if (!p) return;
    rr_device_release(p);
    delete p;
}

extern "C" int rr_plan_get_info(const rr_plan *p, rr_plan_info *info) {
    if (!p || !info) { rr_set_error("null argument"); return 100; }
    std::memset(info, 0, sizeof(*info));
    info->n = p->n;
    info->n_edges = p->n_edges;
    info->n_blocks = p->n_blocks;
    info->n_export = p->n_export;
    info->n_internal_edges = p->n_internal;
    info->max_skew = p->max_skew;
    info->max_indegree = p->max_deg;
    info->max_block_level = p->max_level;
    info->n_outlets_lo = (int32_t)p->n_outlets;
    info->n_dep_edges = (int64_t)p->dep_idx.size();
    info->device_bytes = 0;
    info->renumbered = p->perm.empty() ? 0 : 1;
    info->reach_depth = p->reach_depth;
    info->all_fast = p->all_fast ? 1 : 0;
    info->narrow_blocks = 0;
    for (const rr_blk_meta &m : p->meta) info->narrow_blocks += (m.int_mask & RR_META_NARROW) ? 1 : 0;
    info->n_headwaters = p->n_hw;
    info->n_work = p->n_work;
    return 0;
}

extern "C" int rr_plan_set_coefficients(rr_plan *p, const double *c1, const double *c2, const double *c3,
                                        const double *c4_dt) {
    if (!p || !c1 || !c2 || !c3) { rr_set_error("null argument"); return 100; }
    auto put = [&](std::vector<double> &dst, const double *src) {
        dst.resize(p->n_work);
        if (p->perm.empty()) std::copy(src, src + p->n, dst.begin());
        else for (int64_t k = 0; k < p->n_work; ++k) dst[k] = p->perm[k] >= 0 ? src[p->perm[k]] : 0.0;
    };
    put(p->c1, c1); put(p->c2, c2); put(p->c3, c3);
    p->have_c4 = c4_dt != nullptr;
    if (c4_dt) put(p->c4, c4_dt); else p->c4.assign(p->n_work, 0.0);
    p->coeff_version++;
    return 0;
}

extern "C" int rr_plan_get_arrays(const rr_plan *p, const int32_t **up_ptr, const int32_t **up_idx,
                                  const uint8_t **skew, const int32_t **slot_src, const int32_t **export_id,
                                  const int32_t **blk_level, const int32_t **dep_ptr, const int32_t **dep_idx,
                                  const int32_t **exp_span, const int32_t **perm) {
    if (!p) { rr_set_error("null plan"); return 100; }
    if (up_ptr) *up_ptr = p->up_ptr.data();
    if (up_idx) *up_idx = p->up_idx.data();
    if (skew) *skew = p->skew.data();
    if (slot_src) *slot_src = p->slot_src.data();
    if (export_id) *export_id = p->export_id.data();
    if (blk_level) *blk_level = p->blk_level.data();
    if (dep_ptr) *dep_ptr = p->dep_ptr.data();
    if (dep_idx) *dep_idx = p->dep_idx.data();
    if (exp_span) *exp_span = p->exp_span.data();
    if (perm) *perm = p->perm.empty() ? nullptr : p->perm.data();
    return 0;
}

// ------------------------------------------------------------------------------------------
// Ticket schedule.  Item (block b, tile j) gets key = level(b) + j * delta; tickets are handed
// out in key order.  Every dependency of an item has a strictly smaller key:
//   upstream block, same tile      level(b') < level(b)
//   same block, previous tile      delta >= 1
//   exchange-ring reuse (WAR)      the series of reach u lives in a ring of depth
//                                  R_u = span_u / delta + 1 tiles (span_u = level(consumer) - level(producer));
//                                  producer (b', j) overwrites what consumer (c, j - R_u) read, and
//                                  key(c, j - R_u) < key(b', j)  <=>  span_u < R_u * delta.
// so the warp holding the lowest unfinished ticket can always finish: no deadlock, for any grid
// size, without a cooperative launch.  A small delta lets many tiles be in flight at once (the
// wavefront through deep networks); the budget bounds the ring memory that costs.
// ------------------------------------------------------------------------------------------
static int64_t ring_rows(const rr_plan &p, int64_t delta) {
    int64_t rows = 0;
    for (int32_t sp : p.exp_span) rows += sp / delta + 1;
    return rows;
}

// Exchange rings: depend on the network and the memory budget only (not on the length of the call).
void rr_build_rings(const rr_plan &p, int32_t delta, int64_t budget_rows, rr_schedule &s) {
    if (delta <= 0) {
        int64_t d = 1;
        while (d <= p.max_level && ring_rows(p, d) > budget_rows) d <<= 1;
        delta = (int32_t)std::min<int64_t>(d, (int64_t)p.max_level + 1);
    }
    s.delta = delta;
    s.exp_ro.resize(2 * (size_t)p.n_export);
    int64_t rows = 0;
    for (int64_t e = 0; e < p.n_export; ++e) {
        const int64_t ring = p.exp_span[e] / delta + 1;
        s.exp_ro[2 * e] = (int32_t)rows;
        s.exp_ro[2 * e + 1] = (int32_t)ring;
        rows += ring;
    }
    s.edge_ro.assign(2 * (size_t)p.n_edges, 0);
    for (int64_t e = 0; e < p.n_edges; ++e) {
        const int32_t x = p.export_id[p.up_idx[e]];
        s.edge_ro[2 * e] = x >= 0 ? s.exp_ro[2 * (size_t)x] : -1;
        s.edge_ro[2 * e + 1] = x >= 0 ? s.exp_ro[2 * (size_t)x + 1] : 1;
    }
    s.raw_rows = std::max<int64_t>(rows, 1);
}

// Ticket keys for a call of n_tiles tiles (small: max_level + n_tiles * delta entries).
// blocks of level l that take part in the schedule (all of them, or without the leading first_block level-0 blocks)
static inline int64_t level_count(const rr_plan &p, int32_t l, int64_t first_block) {
    const int64_t c = p.lvl_ptr[l + 1] - p.lvl_ptr[l];
    return l == 0 ? c - first_block : c;
}

void rr_build_keys(const rr_plan &p, int64_t n_tiles, rr_schedule &s, int64_t first_block) {
    s.n_keys = (int64_t)p.max_level + (n_tiles - 1) * (int64_t)s.delta + 1;
    s.key_start.assign(s.n_keys + 1, 0);
    for (int64_t j = 0; j < n_tiles; ++j)
        for (int32_t l = 0; l <= p.max_level; ++l)
            s.key_start[l + j * s.delta + 1] += level_count(p, l, first_block);
    for (int64_t k = 0; k < s.n_keys; ++k) s.key_start[k + 1] += s.key_start[k];
    s.n_items = s.key_start[s.n_keys];
}

// Ticket table: {block, tile, dep_ptr[block], dep_ptr[block + 1]} of every ticket in ticket order -- keys ascending, within a key tiles descending,
// within a (key, tile) the blocks of that level in (level, id) order.  One 16-byte load decodes a ticket on the
// device (the search over key_start that rr_decode_ticket does costs a dozen dependent loads per work item).
void rr_build_items(const rr_plan &p, int64_t n_tiles, const rr_schedule &s, std::vector<int32_t> &items, int64_t first_block) {
    items.resize(4 * (size_t)s.n_items);
    size_t t = 0;
    for (int64_t key = 0; key < s.n_keys; ++key) {
        const int64_t jlo = key > p.max_level ? (key - p.max_level + s.delta - 1) / s.delta : 0;
        const int64_t jhi = std::min<int64_t>(n_tiles - 1, key / s.delta);
        // newest tile (lowest level) first: an item's producers -- the upstream blocks of the same tile and its own
        // previous tile, all in the previous key -- then sit about one whole key of tickets behind it.  (Oldest tile
        // first put the level-1 items of a key right behind their level-0 producers at the end of the previous key: on
        // networks with few blocks per level the consumers were drawn while the producers still ran.)
        for (int64_t j = jhi; j >= jlo; --j) {
            const int64_t l = key - j * s.delta;
            for (int32_t r = p.lvl_ptr[l]; r < p.lvl_ptr[l + 1]; ++r) {
                const int32_t blk = p.lvl_blk[r];
                if (blk < first_block) continue;   // lvl_blk is sorted by (level, id): these are level-0 blocks
                items[4 * t] = blk;
                items[4 * t + 1] = (int32_t)j;
                items[4 * t + 2] = p.dep_ptr[blk];        // the block's upstream-block list, so that the kernel can
                items[4 * t + 3] = p.dep_ptr[blk + 1];    // fetch it without first loading dep_ptr
                ++t;
            }
        }
    }
}

void rr_build_schedule(const rr_plan &p, int64_t n_tiles, int32_t delta, int64_t budget_rows, rr_schedule &s) {
    rr_build_rings(p, delta, budget_rows, s);
    rr_build_keys(p, n_tiles, s);
}

void rr_decode_ticket(const rr_plan &p, const rr_schedule &s, int64_t n_tiles, int64_t ticket,
                      int32_t *block, int32_t *tile) {
    int64_t lo = 0, hi = s.n_keys;  // largest key with key_start[key] <= ticket
    while (hi - lo > 1) {
        const int64_t mid = (lo + hi) >> 1;
        if (s.key_start[mid] <= ticket) lo = mid; else hi = mid;
    }
    int64_t r = ticket - s.key_start[lo];
    const int64_t jlo = lo > p.max_level ? (lo - p.max_level + s.delta - 1) / s.delta : 0;
    const int64_t jhi = std::min<int64_t>(n_tiles - 1, lo / s.delta);
    for (int64_t j = jhi; j >= jlo; --j) {
        const int64_t l = lo - j * s.delta;
        const int64_t w = p.lvl_ptr[l + 1] - p.lvl_ptr[l];
        if (r < w) { *block = p.lvl_blk[p.lvl_ptr[l] + r]; *tile = (int32_t)j; return; }
        r -= w;
    }
    *block = -1; *tile = -1;
}

extern "C" int rr_plan_schedule(const rr_plan *p, int64_t n_tiles, int32_t tile_stride, int64_t *n_items,
                                int32_t *item_block, int32_t *item_tile) {
    if (!p || n_tiles <= 0 || tile_stride <= 0) { rr_set_error("bad argument"); return 100; }
    rr_schedule s;
    rr_build_schedule(*p, n_tiles, tile_stride, INT64_MAX, s);
    if (n_items) *n_items = s.n_items;
    if (item_block && item_tile) {
        // the table the kernel reads, cross-checked against the one-ticket-at-a-time decode
        std::vector<int32_t> items;
        rr_build_items(*p, n_tiles, s, items);
        for (int64_t t = 0; t < s.n_items; ++t) {
            item_block[t] = items[4 * (size_t)t];
            item_tile[t] = items[4 * (size_t)t + 1];
            int32_t b = -1, j = -1;
            rr_decode_ticket(*p, s, n_tiles, t, &b, &j);
            if (b != item_block[t] || j != item_tile[t]) { rr_set_error("ticket table disagrees with the ticket decode"); return 101; }
        }
    }
    return 0;
}
