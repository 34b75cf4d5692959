// Host-only helpers with no dependency on the rest of the library: drainage-basin labelling / bin-packing and
// the synthetic network generator.  Compiled into librr_b200.so and, on its own (-DRR_HOSTUTIL_STANDALONE), into
// oracle/_build/librr_hostutil.so, so that bench.py's reference arm can build the same network and the same
// basin partition without loading the product library.  Reference citations are in include/rr_b200.h.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <numeric>
#include <string>
#include <vector>

#ifdef RR_HOSTUTIL_STANDALONE
static thread_local std::string g_hostutil_err;
static void rr_set_error(const std::string &msg) { g_hostutil_err = msg; }
extern "C" const char *rr_hostutil_last_error(void) { return g_hostutil_err.c_str(); }
#else
#include "rr_internal.h"
#endif

namespace {
inline uint64_t hu_mix64(uint64_t x) {
    x += 0x9e3779b97f4a7c15ull;
    x = (x ^ (x >> 30)) * 0xbf58476d1ce4e5b9ull;
    x = (x ^ (x >> 27)) * 0x94d049bb133111ebull;
    return x ^ (x >> 31);
}
}  // namespace

extern "C" int rr_label_basins(int64_t n, const int32_t *down, int32_t *basin, int64_t *n_basins,
                               int32_t n_parts, int32_t *part) {
    int64_t nb = 0;
    for (int64_t i = 0; i < n; ++i)
        if (down[i] < 0) basin[i] = (int32_t)nb++;
    for (int64_t i = n - 1; i >= 0; --i)
        if (down[i] >= 0) {
            if (down[i] <= i || down[i] >= n) { rr_set_error("down_idx is not topologically sorted"); return 3; }
            basin[i] = basin[down[i]];
        }
    if (n_basins) *n_basins = nb;
    if (part && n_parts > 0) {
        std::vector<int64_t> size(nb, 0);
        for (int64_t i = 0; i < n; ++i) size[basin[i]]++;
        std::vector<int32_t> order(nb);
        std::iota(order.begin(), order.end(), 0);
        std::stable_sort(order.begin(), order.end(), [&](int32_t a, int32_t b) { return size[a] > size[b]; });
        std::vector<int64_t> load(n_parts, 0);
        std::vector<int32_t> owner(nb, 0);
        for (int32_t b : order) {  // LPT greedy: largest basin to the least-loaded part
            int32_t best = 0;
            for (int32_t g = 1; g < n_parts; ++g)
                if (load[g] < load[best]) best = g;
            owner[b] = best;
            load[best] += size[b];
        }
        for (int64_t i = 0; i < n; ++i) part[i] = owner[basin[i]];
    }
    return 0;
}

// ------------------------------------------------------------------------------------------
// Synthetic forests
// ------------------------------------------------------------------------------------------
namespace {
struct Rng {
    uint64_t s[4];
    explicit Rng(uint64_t seed) {
        for (int i = 0; i < 4; ++i) { seed += 0x9e3779b97f4a7c15ull; s[i] = hu_mix64(seed); }
    }
    static inline uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
    uint64_t next() {  // xoshiro256**
        const uint64_t r = rotl(s[1] * 5, 7) * 9, t = s[1] << 17;
        s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = rotl(s[3], 45);
        return r;
    }
    double uniform() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }
    uint64_t below(uint64_t n) { return (uint64_t)(uniform() * (double)n) % n; }
    double normal() {
        double u1 = uniform(), u2 = uniform();
        if (u1 < 1e-300) u1 = 1e-300;
        return std::sqrt(-2.0 * std::log(u1)) * std::cos(6.283185307179586 * u2);
    }
};
}  // namespace

extern "C" int rr_synth_forest(int64_t n, int64_t n_basins, uint64_t seed, double depth_bias,
                               int64_t main_stem, double sigma, int32_t *down) {
    if (n <= 0 || n_basins <= 0 || n_basins > n || n > 0x7ffffff0ll) { rr_set_error("bad forest size"); return 100; }
    Rng rng(seed);
    // basin sizes: lognormal(sigma) normalised to n, each at least 1
    std::vector<double> w(n_basins);
    double tot = 0;
    for (auto &x : w) { x = std::exp(sigma * rng.normal()); tot += x; }
    std::vector<int64_t> size(n_basins);
    int64_t used = 0;
    for (int64_t b = 0; b < n_basins; ++b) {
        size[b] = std::max<int64_t>(1, (int64_t)std::floor(w[b] / tot * (double)(n - n_basins)) + 1);
        used += size[b];
    }
    {   // hand the rounding remainder to the largest basin (or take it back from it)
        int64_t big = std::max_element(size.begin(), size.end()) - size.begin();
        size[big] += n - used;
        if (size[big] < 1) { rr_set_error("basin size normalisation failed"); return 100; }
    }
    if (main_stem > 0) std::swap(size[0], *std::max_element(size.begin(), size.end()));
    std::vector<int32_t> parent, tips;
    std::vector<uint8_t> half;  // tip may take only one more child (pre-seeded stem)
    int64_t off = 0;
    for (int64_t b = 0; b < n_basins; ++b) {
        const int64_t m = size[b];
        parent.assign(m, -1);
        half.assign(m, 0);
        tips.clear();
        int64_t cnt = 1;
        const int64_t stem = (b == 0 && main_stem > 0) ? std::min<int64_t>(main_stem, m) : 0;
        if (stem > 1) {
            for (int64_t g = 1; g < stem; ++g) { parent[g] = (int32_t)(g - 1); half[g - 1] = 1; tips.push_back((int32_t)(g - 1)); }
            tips.push_back((int32_t)(stem - 1));
            cnt = stem;
        } else tips.push_back(0);
        while (cnt < m) {
            size_t pick = tips.size() - 1;
            if (!(rng.uniform() < depth_bias)) pick = (size_t)rng.below(tips.size());
            const int32_t t = tips[pick];
            tips[pick] = tips.back();
            tips.pop_back();
            int64_t kids = (half[t] || rng.uniform() >= 0.7) ? 1 : 2;
            kids = std::min<int64_t>(kids, m - cnt);
            for (int64_t c = 0; c < kids; ++c) { parent[cnt] = t; tips.push_back((int32_t)cnt); cnt++; }
        }
        for (int64_t g = 0; g < m; ++g) {
            const int64_t idx = off + (m - 1 - g);
            down[idx] = parent[g] < 0 ? -1 : (int32_t)(off + (m - 1 - parent[g]));
        }
        off += m;
    }
    return 0;
}
