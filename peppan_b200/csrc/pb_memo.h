// Pair-alignment memo of the clustering path (internal).  iterClust (PEPPAN.py:1777-1792) calls getClust eleven times on
// nested gene sets, so the same (gene, gene, window) alignments come up again at every rung; the alignment of a window
// does not depend on the rung's thresholds, only the decision taken on it does.  The memo remembers, per context, the
// alignment (score, coordinates, match / gap counts) of every window pb_cluster has verified, keyed by the CONTENT of the
// two sequences (64-bit hashes) and the window, and pb_search (cluster mode) skips windows it already knows.  Bounded
// size (PB_CLUSTER_MEMO_MB, default 4096; 0 disables); when full, new results are simply not remembered.
#pragma once
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>

struct MemoVal { int32_t score, qs, qe, ts, te, nm, nx, ngo, ngb; };   // qs == -1: forward result only (failed the pre-trace cuts)

class PairMemo {
public:
    struct Entry { uint64_t k1, k2; MemoVal v; uint32_t used; };
    explicit PairMemo(size_t max_bytes) {
        size_t n = 1;
        while (n * 2 * sizeof(Entry) <= max_bytes) n *= 2;
        cap_ = n; limit_ = n / 10 * 7;
    }
    static uint64_t mix(uint64_t x) { x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33; return x; }
    static void key(uint64_t hq, uint64_t ht, int64_t tbeg, int64_t tlen, uint64_t* k1, uint64_t* k2) {
        *k1 = mix(hq ^ mix(ht + 0x9e3779b97f4a7c15ull)); *k2 = mix(ht ^ mix((uint64_t)tbeg * 0x100000001b3ull + (uint64_t)tlen) ^ (hq << 1));
    }
    const MemoVal* find(uint64_t k1, uint64_t k2) const {
        if (tab_.empty()) return nullptr;
        for (size_t i = k1 & (cap_ - 1);; i = (i + 1) & (cap_ - 1)) {
            const Entry& e = tab_[i];
            if (!e.used) return nullptr;
            if (e.k1 == k1 && e.k2 == k2) return &e.v;
        }
    }
    void put(uint64_t k1, uint64_t k2, const MemoVal& v) {
        if (tab_.empty()) tab_.assign(cap_, Entry{0, 0, {}, 0});      // allocated on first use
        if (size_ >= limit_) { ++dropped_; return; }
        for (size_t i = k1 & (cap_ - 1);; i = (i + 1) & (cap_ - 1)) {
            Entry& e = tab_[i];
            if (!e.used) { e.k1 = k1; e.k2 = k2; e.v = v; e.used = 1; ++size_; return; }
            if (e.k1 == k1 && e.k2 == k2) { e.v = v; return; }
        }
    }
    void clear() { std::vector<Entry>().swap(tab_); size_ = 0; dropped_ = 0; }
    size_t size() const { return size_; }
    size_t dropped() const { return dropped_; }
private:
    std::vector<Entry> tab_;
    size_t cap_ = 0, limit_ = 0, size_ = 0, dropped_ = 0;
};

// FNV-1a over the bytes of a sequence, finalised: the content key of a gene
inline uint64_t pb_seq_hash(const uint8_t* s, int64_t n) {
    uint64_t h = 0xcbf29ce484222325ull;
    for (int64_t i = 0; i < n; ++i) { h ^= s[i]; h *= 0x100000001b3ull; }
    return PairMemo::mix(h ^ (uint64_t)n);
}
