// dto_host.cpp -- host layer mirroring the reference crate's public surface above the CUDA engine:
// collections (RankedFeatureList::from, thresholds), readers, compute_population_size, the string-id ->
// slot canonicalisation, run_single_node's task sharding (threads/MPI ranks -> GPUs), and the exact host
// epilogue (empirical_pvalue, fdr, serde_json-style pretty printing).
// Compile with -ffp-contract=off: generate_thresholds and fdr must round like rustc's f64 code.
#include <algorithm>
#include <charconv>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <atomic>
#include <condition_variable>
#include <deque>
#include <fstream>
#include <memory>
#include <mutex>
#include <numeric>
#include <string>
#include <thread>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include "dto_host_math.hpp"
#include "dto_internal.hpp"

using dto::fail;

struct dto_b200_ranked_list {
    std::vector<std::string> ids;      // sorted by rank (stable)
    std::vector<uint32_t> ranks;       // ascending
    std::vector<uint32_t> thresholds;  // src/collections/ranked.rs:359-375
    uint64_t uid = 0;                  // unique per created list (lists are immutable after creation)
    // id index built once at creation: 64-bit hash per id and an open-addressing table of positions, so that joining
    // two lists (the integer form of intersect_genes) is one flat probe per id instead of a string hash map per pair
    std::vector<uint64_t> id_hash;
    std::vector<int32_t> id_table;     // size = power of two >= 2 n; -1 = empty
    int32_t dup_index = -1;            // position of the first id that repeats an earlier one (-1: all distinct)
    // memo of the last string-id -> slot canonicalisation against another list (a pure function of the two immutable
    // lists); guarded by a mutex because one list may be loaded into several contexts from several threads
    mutable std::mutex memo_mu;
    mutable uint64_t memo_partner_uid = 0;
    mutable std::vector<int32_t> memo_slot;
};

struct dto_b200_feature_list {
    std::vector<std::string> ids;
};

namespace {

std::atomic<uint64_t> g_next_list_uid{1};

// Process-wide pool of idle engine contexts, per device: run_single_node / run_pairs take their contexts from here, so
// only the first call on a device pays for stream, event and buffer creation (~25 ms per context) and repeated calls
// reuse warm buffers.  Contexts are never shared between threads while in use; idle ones are kept until process exit.
std::mutex g_pool_mu;
std::vector<dto_b200_ctx *> g_pool[64];
constexpr size_t kPoolPerDevice = 4;

int pool_acquire(dto_b200_ctx **ctx_out, int device) {
    {
        std::lock_guard<std::mutex> lock(g_pool_mu);
        if (device >= 0 && device < 64 && !g_pool[device].empty()) {
            *ctx_out = g_pool[device].back();
            g_pool[device].pop_back();
            return DTO_B200_OK;
        }
    }
    return dto_b200_create(ctx_out, device);
}

void pool_release(dto_b200_ctx *ctx, int device, bool healthy) {
    if (!ctx) return;
    if (healthy && device >= 0 && device < 64) {
        // an idle context keeps its batch buffers up to 6 GiB each (one 100 000-permutation batch of partner-slot rows at
        // N = 20 000 is 4 GB; re-allocating it on every call costs milliseconds), the pool holds at most kPoolPerDevice
        dto::trim_batch_buffers(ctx, (size_t)6 << 30);
        std::lock_guard<std::mutex> lock(g_pool_mu);
        if (g_pool[device].size() < kPoolPerDevice) {
            g_pool[device].push_back(ctx);
            return;
        }
    }
    dto_b200_destroy(ctx);
}

inline uint64_t hash_id(const std::string &s) {  // FNV-1a, finalised so that the low bits are well mixed
    uint64_t h = 1469598103934665603ull;
    for (unsigned char c : s) h = (h ^ c) * 1099511628211ull;
    h ^= h >> 29;
    h *= 0xBF58476D1CE4E5B9ull;
    h ^= h >> 32;
    return h;
}

void build_id_index(dto_b200_ranked_list *l) {
    const size_t n = l->ids.size();
    size_t cap = 16;
    while (cap < 2 * n) cap <<= 1;
    l->id_hash.resize(n);
    l->id_table.assign(cap, -1);
    l->dup_index = -1;
    for (size_t j = 0; j < n; ++j) {
        const uint64_t h = hash_id(l->ids[j]);
        l->id_hash[j] = h;
        size_t x = (size_t)h & (cap - 1);
        bool dup = false;
        while (l->id_table[x] >= 0) {
            const int32_t o = l->id_table[x];
            if (l->id_hash[(size_t)o] == h && l->ids[(size_t)o] == l->ids[j]) {
                dup = true;
                break;
            }
            x = (x + 1) & (cap - 1);
        }
        if (dup) {
            if (l->dup_index < 0) l->dup_index = (int32_t)j;
        } else {
            l->id_table[x] = (int32_t)j;
        }
    }
}

inline int32_t find_id(const dto_b200_ranked_list *l, const std::string &id, uint64_t h) {
    const size_t cap = l->id_table.size();
    size_t x = (size_t)h & (cap - 1);
    while (l->id_table[x] >= 0) {
        const int32_t o = l->id_table[x];
        if (l->id_hash[(size_t)o] == h && l->ids[(size_t)o] == id) return o;
        x = (x + 1) & (cap - 1);
    }
    return -1;
}

// ln_factorial tables for the host-side re-evaluations of the epilogue, kept per population (building one costs ~1 ms
// per 20 000 entries); a handful of populations at most are alive in one process (one per list pair shape)
std::mutex g_lf_mu;
std::vector<std::pair<uint64_t, std::shared_ptr<std::vector<double>>>> g_lf_cache;

std::shared_ptr<std::vector<double>> host_lf_table(uint64_t N) {
    {
        std::lock_guard<std::mutex> lock(g_lf_mu);
        for (auto &e : g_lf_cache)
            if (e.first == N) return e.second;
    }
    auto t = std::make_shared<std::vector<double>>();
    dto::host_fill_ln_factorial(*t, N);
    std::lock_guard<std::mutex> lock(g_lf_mu);
    if (g_lf_cache.size() >= 4) g_lf_cache.erase(g_lf_cache.begin());
    g_lf_cache.emplace_back(N, t);
    return t;
}

std::string trim_ws(const std::string &s) {
    size_t b = 0, e = s.size();
    auto ws = [](unsigned char c) { return c == ' ' || (c >= 9 && c <= 13); };
    while (b < e && ws((unsigned char)s[b])) ++b;
    while (e > b && ws((unsigned char)s[e - 1])) --e;
    return s.substr(b, e - b);
}

// BufRead::lines(): split on '\n', drop one trailing '\r'
bool read_lines(const char *path, std::vector<std::string> &out) {
    std::ifstream f(path, std::ios::binary);
    if (!f) return false;
    std::string line;
    while (std::getline(f, line)) {
        if (!line.empty() && line.back() == '\r') line.pop_back();
        out.push_back(line);
    }
    return true;
}

// generate_thresholds (src/collections/ranked.rs:359-375).  T1 = 1, T_n = floor(T_{n-1} * 1.01 + 1) while
// <= max rank.  The reference's "set the final threshold to the maximum rank" statement mutates the previous
// (empty) vector and is a no-op, so the series is NOT closed with max_rank; reproduced on purpose.
void generate_thresholds(const std::vector<uint32_t> &sorted_ranks, std::vector<uint32_t> &out) {
    out.clear();
    const uint32_t max_rank = sorted_ranks.empty() ? 0u : sorted_ranks.back();
    uint32_t current = 1;
    while (current <= max_rank) {
        out.push_back(current);
        const double next = std::floor((double)current * 1.01 + 1.0);
        if (next >= 4294967295.0) break;  // `as u32` saturates; the reference would spin forever here
        current = (uint32_t)next;
    }
}

// serde_json prints f64 with ryu: shortest round-trip digits, decimal notation for 1e-5 <= |v| < 1e16
std::string format_f64(double v) {
    if (!std::isfinite(v)) return "null";  // serde_json maps non-finite floats to null
    if (v == 0.0) return std::signbit(v) ? "-0.0" : "0.0";
    char buf[64];
    auto res = std::to_chars(buf, buf + sizeof(buf), v, std::chars_format::scientific);
    std::string s(buf, res.ptr);
    std::string out;
    size_t pos = 0;
    if (s[0] == '-') {
        out = "-";
        pos = 1;
    }
    const size_t epos = s.find('e', pos);
    std::string mant = s.substr(pos, epos - pos);
    const int e10 = std::stoi(s.substr(epos + 1));
    std::string digits;
    for (char c : mant)
        if (c != '.') digits.push_back(c);
    const int length = (int)digits.size();
    const int k = e10 - (length - 1);
    const int kk = length + k;
    if (k >= 0 && kk <= 16) {
        out += digits + std::string((size_t)k, '0') + ".0";
    } else if (kk > 0 && kk <= 16) {
        out += digits.substr(0, (size_t)kk) + "." + digits.substr((size_t)kk);
    } else if (kk > -5 && kk <= 0) {
        out += "0." + std::string((size_t)(-kk), '0') + digits;
    } else if (length == 1) {
        out += digits + "e" + std::to_string(kk - 1);
    } else {
        out += digits.substr(0, 1) + "." + digits.substr(1) + "e" + std::to_string(kk - 1);
    }
    return out;
}

}  // namespace

extern "C" {

// ---------------------------------------------------------------------------------------------------
// collections
// ---------------------------------------------------------------------------------------------------
int dto_b200_ranked_list_from(const char *const *ids, const uint32_t *ranks, size_t n, dto_b200_ranked_list **out) {
    if (!out) return fail(DTO_B200_ERR_INVALID, "null out");
    *out = nullptr;
    if (n && (!ids || !ranks)) return fail(DTO_B200_ERR_INVALID, "null ids/ranks");
    auto *l = new dto_b200_ranked_list();
    l->uid = g_next_list_uid.fetch_add(1);
    // sort_genes_and_ranks (ranked.rs:527-542): stable sort_by_key on rank
    std::vector<uint32_t> order(n);
    std::iota(order.begin(), order.end(), 0u);
    std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return ranks[a] < ranks[b]; });
    l->ids.reserve(n);
    l->ranks.reserve(n);
    for (uint32_t o : order) {
        if (!ids[o]) {
            delete l;
            return fail(DTO_B200_ERR_INVALID, "null feature id at index %u", o);
        }
        l->ids.emplace_back(ids[o]);
        l->ranks.push_back(ranks[o]);
    }
    generate_thresholds(l->ranks, l->thresholds);
    build_id_index(l);
    *out = l;
    return DTO_B200_OK;
}

int dto_b200_read_ranked_list_csv(const char *path, dto_b200_ranked_list **out) {
    if (!path || !out) return fail(DTO_B200_ERR_INVALID, "null argument");
    *out = nullptr;
    std::vector<std::string> lines;
    if (!read_lines(path, lines)) return fail(DTO_B200_ERR_IO, "Could not open file: %s", path);
    std::vector<std::string> ids;
    std::vector<uint32_t> ranks;
    for (size_t i = 0; i < lines.size(); ++i) {
        const std::string &line = lines[i];
        // fields = line.split(','); exactly two fields or panic (read_ranked_feature_list_from_csv.rs:58-61)
        const size_t c = line.find(',');
        if (c == std::string::npos || line.find(',', c + 1) != std::string::npos)
            return fail(DTO_B200_ERR_PANIC, "Invalid format in file %s at line %zu", path, i + 1);
        const std::string f = trim_ws(line.substr(0, c));
        std::string r = trim_ws(line.substr(c + 1));
        // usize::from_str: optional '+', then decimal digits only
        size_t b = (!r.empty() && r[0] == '+') ? 1 : 0;
        if (b >= r.size()) return fail(DTO_B200_ERR_PANIC, "Invalid rank value: \"%s\" (%s line %zu)", r.c_str(), path, i + 1);
        unsigned long long v = 0;
        for (size_t x = b; x < r.size(); ++x) {
            if (r[x] < '0' || r[x] > '9')
                return fail(DTO_B200_ERR_PANIC, "Invalid rank value: \"%s\" (%s line %zu)", r.c_str(), path, i + 1);
            if (v > (0xFFFFFFFFFFFFFFFFull - (unsigned)(r[x] - '0')) / 10ull)
                return fail(DTO_B200_ERR_PANIC, "Invalid rank value: \"%s\" (%s line %zu)", r.c_str(), path, i + 1);
            v = v * 10ull + (unsigned)(r[x] - '0');
        }
        ids.push_back(f);
        ranks.push_back((uint32_t)v);  // `rank as u32` truncates (:65)
    }
    std::vector<const char *> cids(ids.size());
    for (size_t i = 0; i < ids.size(); ++i) cids[i] = ids[i].c_str();
    return dto_b200_ranked_list_from(cids.data(), ranks.data(), ids.size(), out);
}

void dto_b200_ranked_list_free(dto_b200_ranked_list *l) { delete l; }
size_t dto_b200_ranked_list_len(const dto_b200_ranked_list *l) { return l ? l->ids.size() : 0; }
size_t dto_b200_ranked_list_num_thresholds(const dto_b200_ranked_list *l) { return l ? l->thresholds.size() : 0; }
const uint32_t *dto_b200_ranked_list_thresholds(const dto_b200_ranked_list *l) { return l ? l->thresholds.data() : nullptr; }
const uint32_t *dto_b200_ranked_list_ranks(const dto_b200_ranked_list *l) { return l ? l->ranks.data() : nullptr; }
const char *dto_b200_ranked_list_id(const dto_b200_ranked_list *l, size_t i) {
    return (l && i < l->ids.size()) ? l->ids[i].c_str() : nullptr;
}

int dto_b200_feature_list_from(const char *const *ids, size_t n, dto_b200_feature_list **out) {
    if (!out || (n && !ids)) return fail(DTO_B200_ERR_INVALID, "null argument");
    auto *l = new dto_b200_feature_list();
    for (size_t i = 0; i < n; ++i) l->ids.emplace_back(ids[i] ? ids[i] : "");
    *out = l;
    return DTO_B200_OK;
}

int dto_b200_read_feature_list(const char *path, dto_b200_feature_list **out) {
    if (!path || !out) return fail(DTO_B200_ERR_INVALID, "null argument");
    *out = nullptr;
    std::vector<std::string> lines;
    if (!read_lines(path, lines)) return fail(DTO_B200_ERR_IO, "Could not open file: %s", path);
    auto *l = new dto_b200_feature_list();
    for (auto &s : lines) l->ids.push_back(trim_ws(s));  // every line counts, blank ones too (:48-52)
    *out = l;
    return DTO_B200_OK;
}

void dto_b200_feature_list_free(dto_b200_feature_list *l) { delete l; }
size_t dto_b200_feature_list_len(const dto_b200_feature_list *l) { return l ? l->ids.size() : 0; }
const char *dto_b200_feature_list_id(const dto_b200_feature_list *l, size_t i) {
    return (l && i < l->ids.size()) ? l->ids[i].c_str() : nullptr;
}

// ---------------------------------------------------------------------------------------------------
// compute_population_size (src/dto/compute_population_size.rs:66-104)
// ---------------------------------------------------------------------------------------------------
int dto_b200_compute_population_size(const dto_b200_ranked_list *l1, const dto_b200_ranked_list *l2,
                                     const dto_b200_feature_list *background, uint64_t *population_out) {
    if (!l1 || !l2 || !population_out) return fail(DTO_B200_ERR_INVALID, "null argument");
    if (!background) {
        size_t inter = 0;  // FeatureList::intersect keeps list-1 items whose id is in list 2 (feature_list.rs:278-286)
        for (size_t a = 0; a < l1->ids.size(); ++a) inter += find_id(l2, l1->ids[a], l1->id_hash[a]) >= 0 ? 1 : 0;
        if (inter != l1->ids.size() || inter != l2->ids.size())
            return fail(DTO_B200_ERR_PANIC, "If no background is provided, the feature lists must have identical genes.");
        *population_out = inter;
        return DTO_B200_OK;
    }
    std::unordered_set<std::string> bg(background->ids.begin(), background->ids.end());
    const dto_b200_ranked_list *lists[2] = {l1, l2};
    const char *names[2] = {"first", "second"};
    for (int w = 0; w < 2; ++w) {
        std::string missing;
        for (auto &g : lists[w]->ids)
            if (!bg.count(g)) {
                if (!missing.empty()) missing += ", ";
                missing += "\"" + g + "\"";
                if (missing.size() > 600) {
                    missing += ", ...";
                    break;
                }
            }
        if (!missing.empty())
            return fail(DTO_B200_ERR_PANIC, "The following genes in the %s ranked feature list are not in the background: [%s]",
                        names[w], missing.c_str());
    }
    *population_out = background->ids.size();  // lines are counted as-is, duplicates and blanks included (:101)
    return DTO_B200_OK;
}

// string ids -> slot map (the integer form of intersect_genes.rs:38-56): slot[a] = sorted slot in l2 of the gene at sorted
// slot a of l1, or -1.  Memoised per (l1, l2).
static int slot_map_of(const dto_b200_ranked_list *l1, const dto_b200_ranked_list *l2, std::vector<int32_t> &slot) {
    {
        std::lock_guard<std::mutex> lock(l1->memo_mu);
        if (l1->memo_partner_uid == l2->uid && l1->memo_slot.size() == l1->ids.size() && !l1->ids.empty()) {
            slot = l1->memo_slot;
            return DTO_B200_OK;
        }
    }
    if (l2->dup_index >= 0)
        return fail(DTO_B200_ERR_INVALID, "duplicate feature id \"%s\" in the second ranked list",
                    l2->ids[(size_t)l2->dup_index].c_str());
    if (l1->dup_index >= 0)
        return fail(DTO_B200_ERR_INVALID, "duplicate feature id \"%s\" in the first ranked list",
                    l1->ids[(size_t)l1->dup_index].c_str());
    slot.assign(l1->ids.size(), -1);
    for (size_t a = 0; a < l1->ids.size(); ++a) slot[a] = find_id(l2, l1->ids[a], l1->id_hash[a]);
    std::lock_guard<std::mutex> lock(l1->memo_mu);
    l1->memo_partner_uid = l2->uid;
    l1->memo_slot = slot;
    return DTO_B200_OK;
}

int dto_b200_load_lists(dto_b200_ctx *ctx, const dto_b200_ranked_list *l1, const dto_b200_ranked_list *l2,
                        uint64_t population) {
    if (!ctx || !l1 || !l2) return fail(DTO_B200_ERR_INVALID, "null argument");
    std::vector<int32_t> slot;
    int rc = slot_map_of(l1, l2, slot);
    if (rc) return rc;
    return dto_b200_set_problem(ctx, l1->ranks.data(), l1->ranks.size(), l1->thresholds.data(), l1->thresholds.size(),
                                l2->ranks.data(), l2->ranks.size(), l2->thresholds.data(), l2->thresholds.size(),
                                slot.data(), population);
}

int dto_b200_optimize(dto_b200_ctx *ctx, const dto_b200_ranked_list *l1, const dto_b200_ranked_list *l2, int permute,
                      uint64_t population, uint64_t seed, uint64_t perm_id, dto_b200_record *record_out) {
    if (!record_out) return fail(DTO_B200_ERR_INVALID, "null record_out");
    int rc = dto_b200_load_lists(ctx, l1, l2, population);
    if (rc) return rc;
    if (!permute) return dto_b200_run_unpermuted(ctx, record_out);
    return dto_b200_run_permuted_philox(ctx, seed, perm_id, 1, record_out, nullptr);
}

// ---------------------------------------------------------------------------------------------------
// run_single_node (src/run/single_node.rs:83-137) -- tasks sharded over GPUs instead of threads / MPI ranks
// ---------------------------------------------------------------------------------------------------
int dto_b200_run_tasks(const dto_b200_ranked_list *l1, const dto_b200_ranked_list *l2, uint64_t population,
                       const uint64_t *task_ids, const uint8_t *task_permute, size_t n_tasks, const int *devices,
                       size_t n_devices, uint64_t seed, dto_b200_record *records_out) {
    if (!l1 || !l2) return fail(DTO_B200_ERR_INVALID, "null list");
    if (n_tasks == 0) return DTO_B200_OK;
    if (!task_permute || !records_out) return fail(DTO_B200_ERR_INVALID, "null argument");
    std::vector<int> devs;
    if (n_devices == 0 || !devices) devs.push_back(0);
    else devs.assign(devices, devices + n_devices);

    // runs of consecutive permuted tasks with consecutive ids, so each maps onto one Philox id range
    struct Run {
        size_t first;    // index into the task arrays
        uint64_t id;     // Philox id of that task
        size_t count;
    };
    std::vector<Run> runs;
    std::vector<size_t> unperm;
    size_t n_perm = 0;
    for (size_t t = 0; t < n_tasks; ++t) {
        if (!task_permute[t]) {
            unperm.push_back(t);
            continue;
        }
        ++n_perm;
        const uint64_t id = task_ids ? task_ids[t] : (uint64_t)t;
        if (!runs.empty() && runs.back().first + runs.back().count == t && runs.back().id + runs.back().count == id) runs.back().count++;
        else runs.push_back({t, id, 1});
    }
    // contiguous, near-equal shards of the permuted tasks per device (multi_node.rs:114-129 chunks by rank the same way)
    const size_t G = devs.size();
    std::vector<std::vector<Run>> shard(G);
    {
        const size_t per = (n_perm + G - 1) / G;
        size_t g = 0, room = per;
        for (Run r : runs) {
            while (r.count) {
                if (room == 0) {
                    ++g;
                    room = per;
                }
                const size_t take = std::min(room, r.count);
                shard[g].push_back({r.first, r.id, take});
                r.first += take;
                r.id += take;
                r.count -= take;
                room -= take;
            }
        }
    }
    std::vector<int> rcs(G, DTO_B200_OK);
    std::vector<std::string> errs(G);
    auto worker = [&](size_t g) {
        const bool has_work = !shard[g].empty() || (g == 0 && !unperm.empty());
        if (!has_work) return;
        dto_b200_ctx *ctx = nullptr;
        int rc = pool_acquire(&ctx, devs[g]);
        if (rc == DTO_B200_OK) rc = dto_b200_load_lists(ctx, l1, l2, population);
        if (rc == DTO_B200_OK && g == 0 && !unperm.empty()) {
            dto_b200_record r;
            rc = dto_b200_run_unpermuted(ctx, &r);
            if (rc == DTO_B200_OK)
                for (size_t t : unperm) records_out[t] = r;
        }
        for (size_t x = 0; rc == DTO_B200_OK && x < shard[g].size(); ++x)
            rc = dto_b200_run_permuted_philox(ctx, seed, shard[g][x].id, shard[g][x].count, records_out + shard[g][x].first, nullptr);
        if (rc != DTO_B200_OK) errs[g] = dto_b200_last_error();
        rcs[g] = rc;
        pool_release(ctx, devs[g], rc == DTO_B200_OK);
    };
    if (G == 1) {
        worker(0);
    } else {
        std::vector<std::thread> th;
        for (size_t g = 0; g < G; ++g) th.emplace_back(worker, g);
        for (auto &t : th) t.join();
    }
    for (size_t g = 0; g < G; ++g)
        if (rcs[g] != DTO_B200_OK) return fail(rcs[g], "device %d: %s", devs[g], errs[g].c_str());
    return DTO_B200_OK;
}

int dto_b200_run_single_node(const dto_b200_ranked_list *l1, const dto_b200_ranked_list *l2, uint64_t population,
                             const uint8_t *task_permute, size_t n_tasks, const int *devices, size_t n_devices,
                             uint64_t seed, dto_b200_record *records_out) {
    return dto_b200_run_tasks(l1, l2, population, nullptr, task_permute, n_tasks, devices, n_devices, seed, records_out);
}

// ---------------------------------------------------------------------------------------------------
// batched list pairs: the CLI run of src/main.rs:91-165 once per pair, pairs sharded over GPUs
// ---------------------------------------------------------------------------------------------------
namespace {

inline uint64_t pair_seed(uint64_t seed, size_t q) { return seed + (uint64_t)q * 0x9E3779B97F4A7C15ull; }

// two list pairs whose lists carry the same ranks (hence thresholds, set sizes, screen tables) and whose two lists hold the
// same genes can share one launch: only the gene map of the unpermuted task differs
bool same_ranks(const dto_b200_ranked_list *a, const dto_b200_ranked_list *b) {
    return a->ranks.size() == b->ranks.size() && !memcmp(a->ranks.data(), b->ranks.data(), a->ranks.size() * 4);
}

struct PairGroup {
    size_t lo = 0, hi = 0;           // pairs [lo, hi)
    bool batched = false;            // one launch for the whole group (else pair by pair)
    std::vector<int32_t> slot_maps;  // batched: (hi - lo) x n1
    std::vector<uint64_t> seeds;
    int rc = DTO_B200_OK;
    std::string err;
};

}  // namespace

int dto_b200_run_pairs(const dto_b200_ranked_list *const *lists1, const dto_b200_ranked_list *const *lists2,
                       const uint64_t *populations, size_t n_pairs, size_t permutations, const int *devices,
                       size_t n_devices, uint64_t seed, dto_b200_final_result *results_out) {
    if (n_pairs == 0) return DTO_B200_OK;
    if (!lists1 || !lists2 || !populations || !results_out) return fail(DTO_B200_ERR_INVALID, "null argument");
    for (size_t q = 0; q < n_pairs; ++q)
        if (!lists1[q] || !lists2[q]) return fail(DTO_B200_ERR_INVALID, "null list in pair %zu", q);
    std::vector<int> devs;
    if (n_devices == 0 || !devices) devs.push_back(0);
    else devs.assign(devices, devices + n_devices);
    // One context (one host thread driving the GPU) per device.  A 1 000-permutation pair fills a quarter of one wave of
    // scan warps, so consecutive pairs that share their rank structure -- the common case: every list ranks the same
    // genes 1..n -- go through ONE launch of each kernel (dto::run_pair_group); a second host thread per device prepares
    // the next group (string ids -> gene maps) while the GPU works on the current one.  Results are per-pair pure
    // functions of (seed, pair index), so neither the grouping nor the device list changes them.
    const size_t G = devs.size();
    const size_t per = (n_pairs + G - 1) / G;
    std::vector<int> rcs(G, DTO_B200_OK);
    std::vector<std::string> errs(G);
    auto worker = [&](size_t g) {
        const size_t lo = std::min(g * per, n_pairs), hi = std::min(lo + per, n_pairs);
        if (lo >= hi) return;
        dto_b200_ctx *ctx = nullptr;
        int rc = pool_acquire(&ctx, devs[g]);
        if (rc != DTO_B200_OK) {
            errs[g] = dto_b200_last_error();
            rcs[g] = rc;
            return;
        }
        // group boundaries depend on the launch capacity, which depends on the problem shape: plan with the first pair's
        const size_t cap_tasks = (size_t)1 << 17;
        const size_t max_group = std::min<size_t>(256, std::max<size_t>(1, cap_tasks / (permutations + 1)));
        // producer: builds groups ahead of the GPU (bounded queue of 2)
        std::mutex mu;
        std::condition_variable cv;
        std::deque<std::unique_ptr<PairGroup>> ready;
        bool done = false, abort = false;
        std::thread producer([&]() {
            size_t q = lo;
            while (q < hi) {
                auto grp = std::make_unique<PairGroup>();
                grp->lo = q;
                size_t e = q + 1;
                // identical gene sets <=> every id of l1 is in l2 and the lengths agree (ids are distinct)
                std::vector<int32_t> slot;
                int prc = slot_map_of(lists1[q], lists2[q], slot);
                bool ident = prc == DTO_B200_OK && lists1[q]->ids.size() == lists2[q]->ids.size() &&
                             std::find(slot.begin(), slot.end(), -1) == slot.end() && !slot.empty();
                if (prc != DTO_B200_OK) {
                    grp->rc = prc;
                    grp->err = dto_b200_last_error();
                }
                if (ident) {
                    // keep a group's partner-slot rows (2 B per list-1 position and task) under 6 GiB
                    const size_t row_bytes = 2 * (lists1[q]->ids.size() + 256);
                    const size_t by_memory = std::max<size_t>(1, (((size_t)6 << 30) / row_bytes) / (permutations + 1));
                    const size_t group_cap = std::min(max_group, by_memory);
                    grp->batched = true;
                    grp->slot_maps = slot;
                    grp->seeds.push_back(pair_seed(seed, q));
                    while (e < hi && e - q < group_cap && populations[e] == populations[q] && same_ranks(lists1[e], lists1[q]) &&
                           same_ranks(lists2[e], lists2[q])) {
                        std::vector<int32_t> s2;
                        if (slot_map_of(lists1[e], lists2[e], s2) != DTO_B200_OK) break;  // reported when it leads its own group
                        if (s2.size() != slot.size() || std::find(s2.begin(), s2.end(), -1) != s2.end()) break;
                        grp->slot_maps.insert(grp->slot_maps.end(), s2.begin(), s2.end());
                        grp->seeds.push_back(pair_seed(seed, e));
                        ++e;
                    }
                }
                grp->hi = e;
                q = e;
                std::unique_lock<std::mutex> lock(mu);
                cv.wait(lock, [&] { return ready.size() < 2 || abort; });
                if (abort) return;
                ready.push_back(std::move(grp));
                cv.notify_all();
            }
            std::lock_guard<std::mutex> lock(mu);
            done = true;
            cv.notify_all();
        });
        std::vector<dto_b200_record> recs;
        while (rc == DTO_B200_OK) {
            std::unique_ptr<PairGroup> grp;
            {
                std::unique_lock<std::mutex> lock(mu);
                cv.wait(lock, [&] { return !ready.empty() || done; });
                if (ready.empty()) break;
                grp = std::move(ready.front());
                ready.pop_front();
                cv.notify_all();
            }
            if (grp->rc != DTO_B200_OK) {
                rc = fail(grp->rc, "%s", grp->err.c_str());
                break;
            }
            const size_t n = grp->hi - grp->lo;
            rc = dto_b200_load_lists(ctx, lists1[grp->lo], lists2[grp->lo], populations[grp->lo]);
            if (rc != DTO_B200_OK) break;
            recs.resize(n * (permutations + 1));
            if (grp->batched && dto::identical_gene_sets(ctx)) {
                rc = dto::run_pair_group(ctx, grp->slot_maps.data(), grp->seeds.data(), n, permutations, recs.data());
            } else {  // differing gene sets (background mode): pair by pair (n == 1)
                rc = dto_b200_run_unpermuted(ctx, &recs[0]);
                if (rc == DTO_B200_OK && permutations)
                    rc = dto_b200_run_permuted_philox(ctx, pair_seed(seed, grp->lo), 1, permutations, recs.data() + 1, nullptr);
            }
            for (size_t x = 0; rc == DTO_B200_OK && x < n; ++x)
                rc = dto_b200_empirical_pvalue(recs.data() + x * (permutations + 1), permutations + 1, &results_out[grp->lo + x]);
        }
        {
            std::lock_guard<std::mutex> lock(mu);
            abort = true;
            cv.notify_all();
        }
        producer.join();
        if (rc != DTO_B200_OK) errs[g] = dto_b200_last_error();
        rcs[g] = rc;
        pool_release(ctx, devs[g], rc == DTO_B200_OK);
    };
    if (G == 1) {
        worker(0);
    } else {
        std::vector<std::thread> th;
        for (size_t g = 0; g < G; ++g) th.emplace_back(worker, g);
        for (auto &t : th) t.join();
    }
    for (size_t g = 0; g < G; ++g)
        if (rcs[g] != DTO_B200_OK) return fail(rcs[g], "device %d: %s", devs[g], errs[g].c_str());
    return DTO_B200_OK;
}

// ---------------------------------------------------------------------------------------------------
// epilogue: fdr (src/stat_operations/fdr.rs:29-60), empirical_pvalue (src/stat_operations/empirical_pvalue.rs:109-187)
// ---------------------------------------------------------------------------------------------------
int dto_b200_fdr(uint64_t list1_len, uint64_t list2_len, uint64_t overlap, uint64_t population, double sensitivity,
                 double *fdr_out) {
    if (!fdr_out) return fail(DTO_B200_ERR_INVALID, "null fdr_out");
    if (sensitivity <= 0.0) return fail(DTO_B200_ERR_PANIC, "Sensitivity must be greater than 0.");
    const double df = std::fmax((double)overlap / sensitivity, 0.0);
    const double b = std::fmax((double)list1_len - df, 0.0);
    const double r = std::fmax((double)list2_len - df, 0.0);
    const double num = b * r;
    const double den = (double)population * (double)overlap;
    *fdr_out = den > 0.0 ? num / den : 0.0;
    return DTO_B200_OK;
}

// hypergeometric_pvalue (src/stat_operations/hypergeometric_pvalue.rs:33-50) on the HOST, exactly as the tie resolver and
// the epilogue evaluate it: statrs operation order over the host-built ln-factorial table, host libm exp()
int dto_b200_hypergeometric_pvalue_host(uint64_t N, uint64_t K, uint64_t n, uint64_t k, double *pvalue_out) {
    if (!pvalue_out) return fail(DTO_B200_ERR_INVALID, "null pvalue_out");
    if (K > N || n > N)
        return fail(DTO_B200_ERR_PANIC, "Failed to create hypergeometric distribution: successes %llu / draws %llu > population %llu",
                    (unsigned long long)K, (unsigned long long)n, (unsigned long long)N);
    if (N > ((uint64_t)1 << 27)) return fail(DTO_B200_ERR_UNSUPPORTED, "population exceeds 2^27");
    *pvalue_out = dto::host_hypergeom_pvalue_exact(host_lf_table(N)->data(), N, K, n, k);
    return DTO_B200_OK;
}

int dto_b200_empirical_pvalue(const dto_b200_record *records, size_t n, dto_b200_final_result *out) {
    if (!out || (n && !records)) return fail(DTO_B200_ERR_INVALID, "null argument");
    const dto_b200_record *unperm = nullptr;
    size_t n_unperm = 0, n_perm = 0;
    for (size_t i = 0; i < n; ++i) {
        if (records[i].flags & DTO_B200_FLAG_PERMUTED) {
            ++n_perm;
        } else {
            if (!unperm) unperm = &records[i];
            ++n_unperm;
        }
    }
    if (n_unperm != 1)
        fprintf(stderr, "Warning: Expected exactly one unpermuted result, but found %zu.\n", n_unperm);
    if (!unperm) return fail(DTO_B200_ERR_PANIC, "No unpermuted result found in the provided results.");
    out->rank1 = unperm->rank1;
    out->rank2 = unperm->rank2;
    out->set1_len = unperm->set1_len;
    out->set2_len = unperm->set2_len;
    out->population_size = unperm->population_size;
    out->unpermuted_intersection_size = unperm->intersection_size;
    out->unpermuted_pvalue = unperm->pvalue;
    int rc = dto_b200_fdr(unperm->set1_len, unperm->set2_len, unperm->intersection_size, unperm->population_size, 0.8,
                          &out->fdr);
    if (rc) return rc;
    if (n_perm == 0) {
        out->empirical_pvalue = 1.0;
        return DTO_B200_OK;
    }
    // count of permuted p <= unpermuted p (:160-165).  The reference compares values of ONE evaluator (its libm); here
    // permuted records usually carry device-evaluated p-values.  Wherever the two sides are closer than the ambiguity
    // window (e.g. a permutation reproducing the unpermuted (K, n, k), or its mirror image (n, K, k)), both are
    // re-evaluated on the host in statrs order, so the comparison is the one the reference makes on this machine.
    size_t c = 0;
    const double pu = unperm->pvalue;
    std::shared_ptr<std::vector<double>> lf;
    double pu_host = 0.0;
    for (size_t i = 0; i < n; ++i) {
        const dto_b200_record &r = records[i];
        if (!(r.flags & DTO_B200_FLAG_PERMUTED)) continue;
        if (!dto::host_pvalues_ambiguous(r.pvalue, pu) || r.population_size != unperm->population_size ||
            r.set1_len > r.population_size || r.set2_len > r.population_size) {
            if (r.pvalue <= pu) ++c;
            continue;
        }
        if (!lf) {
            if (unperm->population_size > ((uint64_t)1 << 27) || unperm->set1_len > unperm->population_size ||
                unperm->set2_len > unperm->population_size) {  // not a record this library produced: compare as given
                if (r.pvalue <= pu) ++c;
                continue;
            }
            lf = host_lf_table(unperm->population_size);
            pu_host = dto::host_hypergeom_pvalue_exact(lf->data(), unperm->population_size, unperm->set1_len, unperm->set2_len,
                                                       unperm->intersection_size);
        }
        const double pr = dto::host_hypergeom_pvalue_exact(lf->data(), r.population_size, r.set1_len, r.set2_len, r.intersection_size);
        if (pr <= pu_host) ++c;
    }
    out->empirical_pvalue = (double)c / (double)n_perm;  // no +1 correction (:160-165)
    return DTO_B200_OK;
}

int dto_b200_final_result_json(const dto_b200_final_result *r, char *buf, size_t cap, size_t *len_out) {
    if (!r) return fail(DTO_B200_ERR_INVALID, "null result");
    std::string s = "{\n";
    s += "  \"empirical_pvalue\": " + format_f64(r->empirical_pvalue) + ",\n";
    s += "  \"fdr\": " + format_f64(r->fdr) + ",\n";
    s += "  \"population_size\": " + std::to_string(r->population_size) + ",\n";
    s += "  \"rank1\": " + std::to_string(r->rank1) + ",\n";
    s += "  \"rank2\": " + std::to_string(r->rank2) + ",\n";
    s += "  \"set1_len\": " + std::to_string(r->set1_len) + ",\n";
    s += "  \"set2_len\": " + std::to_string(r->set2_len) + ",\n";
    s += "  \"unpermuted_intersection_size\": " + std::to_string(r->unpermuted_intersection_size) + ",\n";
    s += "  \"unpermuted_pvalue\": " + format_f64(r->unpermuted_pvalue) + "\n";
    s += "}";
    if (len_out) *len_out = s.size();
    if (buf && cap) {
        const size_t ncopy = std::min(cap - 1, s.size());
        memcpy(buf, s.data(), ncopy);
        buf[ncopy] = '\0';
    }
    return DTO_B200_OK;
}

}  // extern "C"
