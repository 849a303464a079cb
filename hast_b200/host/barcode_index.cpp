// barcode_index.cpp -- concurrent interning of barcode strings to dense ids.
//
// The reference keys a std::map<std::string, ...> by the barcode text
// (classify.cpp:50-56) and creates the entry for every read, scoring or not
// (:191,208).  Here the text -> id map lives on the host (sharded open
// addressing under per-shard locks, so all parser threads can intern at once)
// and only the dense id travels to the device.
#include <cstring>
#include <mutex>

#include "host.h"

namespace hasthost {

namespace {
inline uint64_t hash_bytes(const char* s, size_t n) {
    uint64_t h = 0x9E3779B97F4A7C15ull ^ (n * 0xff51afd7ed558ccdull);
    while (n >= 8) {
        uint64_t w;
        memcpy(&w, s, 8);
        h = (h ^ w) * 0xff51afd7ed558ccdull;
        h ^= h >> 32;
        s += 8; n -= 8;
    }
    uint64_t w = 0;
    memcpy(&w, s, n);
    h = (h ^ w) * 0xc4ceb9fe1a85ec53ull;
    h ^= h >> 29;
    h *= 0xff51afd7ed558ccdull;
    h ^= h >> 32;
    return h;
}
struct Entry { uint64_t hash; uint64_t off; uint32_t len; uint32_t id; };
}  // namespace

struct BarcodeIndex::Shard {
    std::mutex mu;
    std::vector<Entry> tab;          // power-of-two open addressing; len == UINT32_MAX marks empty
    std::vector<char> arena;
    size_t used = 0;
    char pad[64];

    Shard() { tab.assign(1024, Entry{0, 0, UINT32_MAX, 0}); }
    void grow() {
        std::vector<Entry> old;
        old.swap(tab);
        tab.assign(old.size() * 2, Entry{0, 0, UINT32_MAX, 0});
        const size_t m = tab.size() - 1;
        for (const Entry& e : old) {
            if (e.len == UINT32_MAX) continue;
            size_t i = (size_t)(e.hash >> 8) & m;
            while (tab[i].len != UINT32_MAX) i = (i + 1) & m;
            tab[i] = e;
        }
    }
};

BarcodeIndex::BarcodeIndex() : shards_(new Shard[kShards]) {}
BarcodeIndex::~BarcodeIndex() = default;

uint32_t BarcodeIndex::intern(const char* s, size_t n) {
    const uint64_t h = hash_bytes(s, n);
    Shard& sh = shards_[h & (kShards - 1)];
    std::lock_guard<std::mutex> lk(sh.mu);
    size_t m = sh.tab.size() - 1;
    size_t i = (size_t)(h >> 8) & m;
    for (;;) {
        Entry& e = sh.tab[i];
        if (e.len == UINT32_MAX) break;
        if (e.hash == h && e.len == n && memcmp(sh.arena.data() + e.off, s, n) == 0) return e.id;
        i = (i + 1) & m;
    }
    const uint32_t id = next_id_.fetch_add(1, std::memory_order_acq_rel);
    const uint64_t off = sh.arena.size();
    sh.arena.insert(sh.arena.end(), s, s + n);
    sh.tab[i] = Entry{h, off, (uint32_t)n, id};
    if (++sh.used * 2 > sh.tab.size()) sh.grow();
    return id;
}

void BarcodeIndex::export_names(std::vector<std::string>& out) const {
    out.assign(size(), std::string());
    for (int s = 0; s < kShards; ++s) {
        const Shard& sh = shards_[s];
        for (const Entry& e : sh.tab)
            if (e.len != UINT32_MAX) out[e.id].assign(sh.arena.data() + e.off, e.len);
    }
}

}  // namespace hasthost
