// barcode_index.cpp -- concurrent interning of barcode strings to dense ids.
//
// The reference keys a std::map<std::string, ...> by the barcode text
// (classify.cpp:50-56) and creates the entry for every read, scoring or not
// (:191,208).  Here the text -> id map lives on the host and only the dense id
// travels to the device.  Interning runs once per read on every parser thread,
// so the common case -- a barcode seen before -- takes no lock: every shard
// publishes an immutable-size open-addressing table through an atomic pointer,
// entries become visible by a release store of their hash, strings live in
// chunks that never move, and a table that fills up is replaced by a larger
// copy while the old one stays alive for the readers still probing it (they
// can only miss, and a miss re-probes the current table under the shard lock).
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <new>
#include <mutex>
#include <sys/mman.h>
#include <thread>

#include "host.h"

namespace hasthost {

namespace {
inline uint64_t load64(const char* s) { uint64_t w; memcpy(&w, s, 8); return w; }
inline uint64_t load32(const char* s) { uint32_t w; memcpy(&w, s, 4); return w; }
inline uint64_t hash_bytes(const char* s, size_t n) {
    // Names of 1..16 bytes (every stLFR barcode, "1234_567_89") take two possibly overlapping fixed-size loads
    // and no loop; the length is mixed in, so a name and its prefix differ.
    uint64_t a, b;
    if (n >= 8 && n <= 16) { a = load64(s); b = load64(s + n - 8); }
    else if (n >= 4 && n < 8) { a = load32(s); b = load32(s + n - 4); }
    else if (n < 4) { a = n ? ((uint64_t)(uint8_t)s[0] | (uint64_t)(uint8_t)s[n >> 1] << 8 | (uint64_t)(uint8_t)s[n - 1] << 16) : 0; b = 0; }
    else {
        uint64_t h = 0x9E3779B97F4A7C15ull ^ (n * 0xff51afd7ed558ccdull);
        size_t m = n;
        const char* p = s;
        while (m > 16) { h = (h ^ load64(p)) * 0xff51afd7ed558ccdull; h ^= h >> 32; p += 8; m -= 8; }
        a = load64(p) ^ h; b = load64(p + m - 8);
    }
    uint64_t h = (a ^ 0x9E3779B97F4A7C15ull ^ (n * 0xff51afd7ed558ccdull)) * 0xff51afd7ed558ccdull;
    h ^= h >> 32;
    h = (h ^ b) * 0xc4ceb9fe1a85ec53ull;
    h ^= h >> 29;
    h *= 0xff51afd7ed558ccdull;
    h ^= h >> 32;
    return h ? h : 1;                 // 0 marks an empty slot
}
// 32 bytes, two per cache line: a lookup of a typical stLFR barcode ("1234_567_89": <= 16 bytes, kept inline)
// touches one line; longer names live in the shard's arena.
struct Entry {
    std::atomic<uint64_t> hash{0};   // written last (release): the other fields are complete when it is non-zero
    uint32_t id = 0;
    uint32_t len = 0;
    union { char text[16]; const char* ptr; } u{};
    const char* data() const { return len <= sizeof(u.text) ? u.text : u.ptr; }
};
static_assert(sizeof(Entry) == 32, "two entries per cache line");
// A run with hundreds of thousands of barcodes probes tables far larger than the TLB reaches with 4 KiB
// pages: every probe -- and every prefetch of one -- then starts with a page walk (measured: the prefetch
// alone cost 67 ns per read, a quarter of the parser).  All tables of an index are therefore carved out of one
// reserved region for which transparent huge pages are requested; pages are committed (and zeroed: an
// all-zero Entry is an empty slot) on first touch, nothing is ever handed back before the index dies.
class HugeArena {
public:
    HugeArena() {
        void* p = mmap(nullptr, kReserve, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
        if (p == MAP_FAILED) return;
        map_ = p;
        base_ = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(p) + kHuge - 1) & ~(uintptr_t)(kHuge - 1));
        cap_ = kReserve - kHuge;
#ifdef MADV_HUGEPAGE
        madvise(base_, cap_, MADV_HUGEPAGE);
#endif
    }
    ~HugeArena() { if (map_) munmap(map_, kReserve); }
    HugeArena(const HugeArena&) = delete;
    HugeArena& operator=(const HugeArena&) = delete;
    void* alloc(size_t bytes) {                            // zeroed, 64-byte aligned; nullptr when exhausted
        bytes = (bytes + 63) & ~(size_t)63;
        const size_t off = top_.fetch_add(bytes, std::memory_order_relaxed);
        return (base_ && off + bytes <= cap_) ? base_ + off : nullptr;
    }
private:
    static constexpr size_t kHuge = (size_t)2 << 20;
    static constexpr size_t kReserve = (size_t)64 << 30;   // address space only
    void* map_ = nullptr;
    char* base_ = nullptr;
    size_t cap_ = 0;
    std::atomic<size_t> top_{0};
};
struct Table {
    Table(size_t n, HugeArena& arena) : mask(n - 1) {
        slots = static_cast<Entry*>(arena.alloc(n * sizeof(Entry)));
        if (!slots) {                                    // only when the arena could not serve
            owned = aligned_alloc(64, n * sizeof(Entry));
            if (!owned) throw std::bad_alloc();
            memset(owned, 0, n * sizeof(Entry));
            slots = static_cast<Entry*>(owned);
        }
    }
    ~Table() { free(owned); }
    Table(const Table&) = delete;
    // what readers load: the slot array (64-byte aligned) with log2 of its size in the low bits, so that a
    // lookup is two dependent loads (this word, then the slot) instead of three
    uintptr_t tagged() const { return reinterpret_cast<uintptr_t>(slots) | (uintptr_t)__builtin_ctzll(mask + 1); }
    size_t mask;
    Entry* slots;
    void* owned = nullptr;
};
inline const Entry* tagged_slots(uintptr_t t) { return reinterpret_cast<const Entry*>(t & ~(uintptr_t)63); }
inline size_t tagged_mask(uintptr_t t) { return ((size_t)1 << (t & 63)) - 1; }
constexpr size_t kArenaChunk = 1u << 16;
}  // namespace

struct BarcodeIndex::Shard {
    std::atomic<Table*> cur{nullptr};
    std::atomic<uintptr_t>* pub = nullptr;               // this shard's word in BarcodeIndex::Arena::pub
    std::mutex mu;
    std::vector<std::unique_ptr<Table>> tables;          // every generation stays alive
    std::vector<std::unique_ptr<char[]>> chunks;         // strings never move
    size_t chunk_used = kArenaChunk;
    size_t used = 0;
    char pad[64];

    HugeArena* arena = nullptr;
    void init(HugeArena* a, std::atomic<uintptr_t>* p) {
        arena = a;
        pub = p;
        tables.emplace_back(new Table(1024, *arena));
        cur.store(tables.back().get(), std::memory_order_release);
        pub->store(tables.back()->tagged(), std::memory_order_release);
    }
    const char* store(const char* s, size_t n) {
        if (n > kArenaChunk / 4) {                       // an oversized name gets a chunk of its own
            chunks.emplace_back(new char[n]);
            memcpy(chunks.back().get(), s, n);
            const char* p = chunks.back().get();
            chunk_used = kArenaChunk;                    // next small string opens a fresh chunk
            return p;
        }
        if (chunk_used + n > kArenaChunk) { chunks.emplace_back(new char[kArenaChunk]); chunk_used = 0; }
        char* p = chunks.back().get() + chunk_used;
        memcpy(p, s, n);
        chunk_used += n;
        return p;
    }
    void grow() {
        Table* old = cur.load(std::memory_order_relaxed);
        std::unique_ptr<Table> nt(new Table((old->mask + 1) * 2, *arena));
        for (size_t i = 0; i <= old->mask; ++i) {
            const Entry& e = old->slots[i];
            const uint64_t h = e.hash.load(std::memory_order_relaxed);
            if (!h) continue;
            size_t j = (size_t)(h >> 8) & nt->mask;
            while (nt->slots[j].hash.load(std::memory_order_relaxed)) j = (j + 1) & nt->mask;
            nt->slots[j].u = e.u;
            nt->slots[j].len = e.len;
            nt->slots[j].id = e.id;
            nt->slots[j].hash.store(h, std::memory_order_relaxed);
        }
        tables.push_back(std::move(nt));
        cur.store(tables.back().get(), std::memory_order_release);
        pub->store(tables.back()->tagged(), std::memory_order_release);
    }
};

struct BarcodeIndex::Arena : HugeArena {
    alignas(64) std::atomic<uintptr_t> pub[BarcodeIndex::kShards];   // current table of every shard, 2 KiB: cache resident
};
BarcodeIndex::BarcodeIndex() : arena_(new Arena()), shards_(new Shard[kShards]) {
    for (int s = 0; s < kShards; ++s) shards_[s].init(arena_.get(), &arena_->pub[s]);
}
BarcodeIndex::~BarcodeIndex() = default;

uint64_t BarcodeIndex::hash(const char* s, size_t n) { return hash_bytes(s, n); }

void BarcodeIndex::prefetch(uint64_t h) const {
    const uintptr_t t = arena_->pub[h & (kShards - 1)].load(std::memory_order_acquire);
    __builtin_prefetch(&tagged_slots(t)[(size_t)(h >> 8) & tagged_mask(t)]);
}

uint32_t BarcodeIndex::intern(const char* s, size_t n) { return intern_hashed(hash_bytes(s, n), s, n); }

uint32_t BarcodeIndex::intern_hashed(uint64_t h, const char* s, size_t n) {
    {                                                    // lock-free: seen before
        const uintptr_t t = arena_->pub[h & (kShards - 1)].load(std::memory_order_acquire);
        const Entry* const slots = tagged_slots(t);
        const size_t mask = tagged_mask(t);
        size_t i = (size_t)(h >> 8) & mask;
        for (;;) {
            const Entry& e = slots[i];
            const uint64_t eh = e.hash.load(std::memory_order_acquire);
            if (!eh) break;
            if (eh == h && e.len == n && memcmp(e.data(), s, n) == 0) return e.id;
            i = (i + 1) & mask;
        }
    }
    Shard& sh = shards_[h & (kShards - 1)];
    std::lock_guard<std::mutex> lk(sh.mu);
    Table* t = sh.cur.load(std::memory_order_relaxed);
    size_t i = (size_t)(h >> 8) & t->mask;
    for (;;) {
        Entry& e = t->slots[i];
        const uint64_t eh = e.hash.load(std::memory_order_relaxed);
        if (!eh) break;
        if (eh == h && e.len == n && memcmp(e.data(), s, n) == 0) return e.id;
        i = (i + 1) & t->mask;
    }
    const uint32_t id = next_id_.fetch_add(1, std::memory_order_acq_rel);
    Entry& e = t->slots[i];
    if (n <= sizeof(e.u.text)) memcpy(e.u.text, s, n); else e.u.ptr = sh.store(s, n);
    e.len = (uint32_t)n;
    e.id = id;
    e.hash.store(h, std::memory_order_release);
    if (++sh.used * 2 > t->mask + 1) sh.grow();
    return id;
}

void BarcodeIndex::export_names(std::vector<std::string>& out) const {
    out.assign(size(), std::string());
    // ids are unique, so the shards can be walked side by side (tens of millions of names at human scale)
    const unsigned threads = out.size() < 200000 ? 1u : std::min(16u, std::max(1u, std::thread::hardware_concurrency()));
    std::atomic<int> next{0};
    auto work = [&]() {
        for (int s; (s = next.fetch_add(1)) < kShards;) {
            const Table* t = shards_[s].cur.load(std::memory_order_acquire);
            for (size_t i = 0; i <= t->mask; ++i) {
                const Entry& e = t->slots[i];
                if (e.hash.load(std::memory_order_acquire)) out[e.id].assign(e.data(), e.len);
            }
        }
    };
    std::vector<std::thread> th;
    for (unsigned i = 1; i < threads; ++i) th.emplace_back(work);
    work();
    for (auto& x : th) x.join();
}

}  // namespace hasthost
