// Host-side building blocks of the coverage-capping read selection that runs before the DP
// (reference: whatshap/readselect.pyx, whatshap/priorityqueue.pyx).
//
// The selection is a sequential greedy heuristic; which reads it picks among equal scores is decided by
// (a) the exact sift rules of the reference's positional binary heap, (b) the iteration order of a
// std::unordered_set<int> of variant positions (libstdc++), and (c) the iteration order of CPython sets
// of read indices.  (a) and (b) live here -- (b) by using the very same container --, (c) stays in
// Python (whatshap_b200/readselect.py) where real `set` objects go through the same operations.
// No CUDA in this file: the step is host work in the reference too.
#include <algorithm>
#include <cstdint>
#include <unordered_set>
#include <vector>

#include "../../include/whmec.h"

namespace {

struct Score {
    int32_t v[3];  // (new - gaps, total - gaps, min quality), readselect.pyx:55-92
};

// priorityqueue.pyx:11-21 for equal-length vectors: strict lexicographic "<"
inline bool lower(const Score &a, const Score &b) {
    for (int i = 0; i < 3; ++i) {
        if (a.v[i] < b.v[i]) return true;
        if (a.v[i] > b.v[i]) return false;
    }
    return false;
}

// Max-heap over (score, item) with item -> slot lookup; sift rules of priorityqueue.pyx:96-127
// (children compared with strict "<": on a tie the LEFT child is the candidate; a parent moves only
// for a strictly larger child).
class ScoreHeap {
public:
    explicit ScoreHeap(uint32_t n_items) : slot_(n_items, -1) {}
    size_t size() const { return heap_.size(); }
    bool contains(uint32_t item) const { return item < slot_.size() && slot_[item] >= 0; }
    const Score &score_of(uint32_t item) const { return heap_[slot_[item]].s; }

    void push(const Score &s, uint32_t item) {
        heap_.push_back({s, item});
        slot_[item] = (int64_t)heap_.size() - 1;
        up(heap_.size() - 1);
    }
    void pop(Score *s, uint32_t *item) {
        const Node top = heap_.front(), last = heap_.back();
        heap_.pop_back();
        slot_[top.item] = -1;
        if (!heap_.empty()) {
            heap_[0] = last;
            slot_[last.item] = 0;
            down(0);
        }
        *s = top.s;
        *item = top.item;
    }
    void change(uint32_t item, const Score &s) {  // priorityqueue.pyx:164-179
        const size_t i = (size_t)slot_[item];
        const Score old = heap_[i].s;
        heap_[i].s = s;
        if (lower(old, s)) up(i);
        else down(i);
    }

private:
    struct Node {
        Score s;
        uint32_t item;
    };
    void swap_nodes(size_t a, size_t b) {
        std::swap(heap_[a], heap_[b]);
        slot_[heap_[a].item] = (int64_t)a;
        slot_[heap_[b].item] = (int64_t)b;
    }
    void up(size_t i) {
        while (i > 0) {
            const size_t p = (i - 1) / 2;
            if (!lower(heap_[p].s, heap_[i].s)) break;
            swap_nodes(p, i);
            i = p;
        }
    }
    void down(size_t i) {
        const size_t n = heap_.size();
        for (;;) {
            const size_t l = 2 * i + 1, r = 2 * i + 2;
            size_t c;
            if (r < n) c = lower(heap_[l].s, heap_[r].s) ? r : l;
            else if (l < n) c = l;
            else return;
            if (!lower(heap_[i].s, heap_[c].s)) return;
            swap_nodes(c, i);
            i = c;
        }
    }
    std::vector<Node> heap_;
    std::vector<int64_t> slot_;
};

// Union-find over variant ranks: which variants the reads taken so far connect (the role of
// whatshap/graph.py:35-86 in readselect.pyx:193-197,206-234; only the partition matters here).
class Blocks {
public:
    void reset(uint32_t n) {
        parent_.resize(n);
        for (uint32_t i = 0; i < n; ++i) parent_[i] = i;
    }
    uint32_t find(uint32_t i) {
        uint32_t root = i;
        while (parent_[root] != root) root = parent_[root];
        while (parent_[i] != root) {
            const uint32_t next = parent_[i];
            parent_[i] = root;
            i = next;
        }
        return root;
    }
    void join(uint32_t a, uint32_t b) {
        a = find(a);
        b = find(b);
        if (a != b) parent_[b] = a;
    }

private:
    std::vector<uint32_t> parent_;
};

struct Selector {
    Selector(uint32_t n_reads, uint32_t n_variants, const uint64_t *read_off, const int32_t *ent_pos, const uint32_t *ent_rank,
             const int32_t *positions, uint32_t max_cov)
        : heap(n_reads), n_variants(n_variants), read_off(read_off), ent_pos(ent_pos), ent_rank(ent_rank), positions(positions),
          max_cov(max_cov), coverage(n_variants, 0), covered(n_variants, 0) {}
    ScoreHeap heap;
    std::unordered_set<int> fresh;  // positions newly covered by the read just popped (readselect.pyx:117,124-137)
    uint32_t n_variants;
    const uint64_t *read_off;   // reads as CSR over entries, ReadSet order (caller keeps the arrays alive)
    const int32_t *ent_pos;     // genomic position of every entry
    const uint32_t *ent_rank;   // its rank among the sorted distinct positions ("vcf index")
    const int32_t *positions;   // sorted distinct positions
    uint32_t max_cov;
    std::vector<uint32_t> coverage;  // whatshap/coverage.py: reads taken per variant rank (physical span)
    std::vector<uint8_t> covered;    // per slice: variant already covered by a read of this slice
    Blocks blocks;

    uint32_t first_rank(uint32_t r) const { return ent_rank[read_off[r]]; }
    uint32_t end_rank(uint32_t r) const { return ent_rank[read_off[r + 1] - 1] + 1; }
    bool full(uint32_t r) const {  // readselect.pyx:141-143
        uint32_t mx = 0;
        for (uint32_t v = first_rank(r); v < end_rank(r); ++v) mx = coverage[v] > mx ? coverage[v] : mx;
        return mx >= max_cov;
    }
    void take(uint32_t r) {
        for (uint32_t v = first_rank(r); v < end_rank(r); ++v) coverage[v] += 1;
    }
    void push_all(const uint32_t *items, const int32_t *scores, uint32_t n) {
        for (uint32_t i = 0; i < n; ++i) {
            const int32_t *sc = scores + 3 * (size_t)items[i];
            heap.push(Score{{sc[0], sc[1], sc[2]}}, items[i]);
        }
    }
};

}  // namespace

extern "C" {

whmec_selector *whmec_selector_create(uint32_t n_reads, uint32_t n_variants, const uint64_t *read_off, const int32_t *ent_pos,
                                      const uint32_t *ent_rank, const int32_t *positions, uint32_t max_cov) {
    return reinterpret_cast<whmec_selector *>(new Selector(n_reads, n_variants, read_off, ent_pos, ent_rank, positions, max_cov));
}

void whmec_selector_destroy(whmec_selector *h) { delete reinterpret_cast<Selector *>(h); }

void whmec_selector_begin_slice(whmec_selector *h, const uint32_t *items, const int32_t *scores, uint32_t n) {
    Selector *s = reinterpret_cast<Selector *>(h);
    std::unordered_set<int>().swap(s->fresh);  // a slice starts with a newly constructed container (readselect.pyx:117)
    std::fill(s->covered.begin(), s->covered.end(), 0);
    s->push_all(items, scores, n);
}

int whmec_selector_next(whmec_selector *h, uint32_t *read, uint32_t *fresh_ranks, uint32_t *n_fresh, uint32_t *over,
                        uint32_t *n_over) {
    Selector *s = reinterpret_cast<Selector *>(h);
    *n_over = 0;
    *n_fresh = 0;
    while (s->heap.size() > 0) {
        Score sc;
        uint32_t r;
        s->heap.pop(&sc, &r);
        // the container is refilled for every popped read, taken or not (its bucket array only ever grows)
        s->fresh.clear();
        for (uint64_t e = s->read_off[r]; e < s->read_off[r + 1]; ++e)
            if (!s->covered[s->ent_rank[e]]) s->fresh.insert(s->ent_pos[e]);
        if (s->full(r)) {
            over[(*n_over)++] = r;
            continue;
        }
        if (s->fresh.empty()) continue;  // covers nothing new: stays undecided for the next slice
        s->take(r);
        uint32_t k = 0;
        for (int p : s->fresh) {  // iteration order of the container = order of the caller's score updates
            const uint32_t v = (uint32_t)(std::lower_bound(s->positions, s->positions + s->n_variants, p) - s->positions);
            s->covered[v] = 1;
            fresh_ranks[k++] = v;
        }
        *n_fresh = k;
        *read = r;
        return 1;
    }
    return 0;
}

void whmec_selector_rescore(whmec_selector *h, const uint32_t *items, uint32_t n) {
    Selector *s = reinterpret_cast<Selector *>(h);
    for (uint32_t i = 0; i < n; ++i) {
        const uint32_t item = items[i];
        if (!s->heap.contains(item)) continue;
        Score sc = s->heap.score_of(item);
        // readselect.pyx:38-52: the first component drops by one for every variant of the read that is NOT
        // among the freshly covered ones
        for (uint64_t e = s->read_off[item]; e < s->read_off[item + 1]; ++e)
            if (s->fresh.find(s->ent_pos[e]) == s->fresh.end()) sc.v[0] -= 1;
        s->heap.change(item, sc);
    }
}

uint32_t whmec_selector_bridge(whmec_selector *h, const uint32_t *items, const int32_t *scores, uint32_t n,
                               const uint32_t *slice_reads, uint32_t n_slice, uint32_t *removed, uint8_t *taken) {
    Selector *s = reinterpret_cast<Selector *>(h);
    s->blocks.reset(s->n_variants);
    auto join_read = [&](uint32_t r) {
        for (uint64_t e = s->read_off[r] + 1; e < s->read_off[r + 1]; ++e) s->blocks.join(s->ent_rank[s->read_off[r]], s->ent_rank[e]);
    };
    for (uint32_t i = 0; i < n_slice; ++i) join_read(slice_reads[i]);
    s->push_all(items, scores, n);
    uint32_t k = 0;
    while (s->heap.size() > 0) {
        Score sc;
        uint32_t r;
        s->heap.pop(&sc, &r);
        if (s->full(r)) {  // can never be added any more
            removed[k] = r;
            taken[k++] = 0;
            continue;
        }
        const uint32_t b0 = s->blocks.find(s->ent_rank[s->read_off[r]]);
        bool bridges = false;
        for (uint64_t e = s->read_off[r] + 1; e < s->read_off[r + 1] && !bridges; ++e) bridges = s->blocks.find(s->ent_rank[e]) != b0;
        if (!bridges) continue;  // inside one block: stays undecided
        s->take(r);
        join_read(r);
        removed[k] = r;
        taken[k++] = 1;
    }
    return k;
}

}  // extern "C"
