// kg_fsm.cuh -- gatherHits / processSetOfHits (KGJ:457-514, 385-455) as a streaming state machine.
//
// The reference keeps the open run as an ArrayList<Hit> and re-scans it in processSetOfHits.  Everything that scan
// computes is a fold over the list in list order with `currentFI` fixed for the lifetime of the list (currentFI only
// changes when the list is empty or inside processSetOfHits), so it can be carried incrementally:
//   n            hits.size()
//   first_pos    hits.get(0).from0InProt
//   cnt, w       fICount and weightedHits: fp32 additions in list order, the very additions KGJ:394 performs
//   last_match   hits.get(lastHit).from0InProt
//   l1, l2       the last two list entries (gap test, order constraint, pair-switch test, retained pair)
//   otu_t        the OTU buffer as it would be after replaying every counted hit of the open run (KGJ:413-438);
//                committed to otu_c when the run is called, dropped otherwise.  Consecutive counted hits with the same
//                OTU index are held back as (pend_oI, pend_cnt) and applied in one step (kg_otu_update_n: m updates of
//                one index collapse exactly), which is what keeps this FSM off the instruction-issue limit
// No list, no second pass, identical results (bit-exact, including the fp32 sum).
#pragma once

#include "kg_common.cuh"

struct KgOtuBuf { // ArrayList<OtuCount>, at most OI_BUFSZ entries (KGJ:99, 1221-1224)
    int n;
    int c[KG_OI_BUFSZ];
    int o[KG_OI_BUFSZ];
};

__device__ __forceinline__ void kg_otu_clear(KgOtuBuf& u) {
    u.n = 0;
#pragma unroll
    for (int i = 0; i < KG_OI_BUFSZ; i++) u.c[i] = u.o[i] = 0;
}

__device__ __forceinline__ void kg_otu_update_n(KgOtuBuf& u, int oI, int m);
// KGJ:415-437: one update.
__device__ __forceinline__ void kg_otu_update(KgOtuBuf& u, int oI) { kg_otu_update_n(u, oI, 1); }

// count += m for oI (m >= 1 consecutive updates of the same OTU index collapse exactly: the entry ends up in front of
// the maximal block of entries ahead of it whose count is <= its final count, whether it got there in one step or m).
//
// The list is always sorted by count, non-increasing: a new or overwritten entry starts at the end and bubbles past every
// entry with a count <= its own (KGJ:432-437), and an incremented entry only ever moves forward.  So the bubbling has a
// closed form -- the entry lands behind the entries whose count is LARGER, the ones between that place and its old one
// shift back by one.  The OTU fold of a long contig is one such update after the other, so what counts is the LENGTH OF
// THE DEPENDENT CHAIN, and the update is written as one-hot selects with no branch and no index arithmetic:
//   e[i]  = entry i holds oI                         (at most one)
//   C     = m + the matched count                    (an OR of five masked counts: one of them at most is non-zero)
//   g[i]  = c[i] > C, i < 4                          (a prefix 1..10..0 because the list is sorted; the matched entry and
//                                                     everything behind it have a count < C, the victim of an overwrite is
//                                                     entry 4 and an appended entry lands on an empty slot whose count is 0,
//                                                     so no "ahead of the old place" test is needed)
//   le[i] = i <= old place j                         (found: no match before i; not found: j = min(n, 4), KGJ:419-426)
//   new[i] = g[i] ? old[i] : (i == 0 || g[i-1]) ? (oI, C) : le[i] ? old[i-1] : old[i]
// About nine dependent operations instead of ~27 (r01: 287 cycles per update in k_otu_fold).
__device__ __forceinline__ void kg_otu_update_n(KgOtuBuf& u, int oI, int m) {
    bool e[KG_OI_BUFSZ];
#pragma unroll
    for (int i = 0; i < KG_OI_BUFSZ; i++) e[i] = i < u.n && u.o[i] == oI;
    const bool found = e[0] | e[1] | e[2] | e[3] | e[4];
    const int C = m + ((e[0] ? u.c[0] : 0) | (e[1] ? u.c[1] : 0) | (e[2] ? u.c[2] : 0) | (e[3] ? u.c[3] : 0) | (e[4] ? u.c[4] : 0));
    bool le[KG_OI_BUFSZ]; // le[0] is always true and never read
    bool before = false;  // a match at an index < i
#pragma unroll
    for (int i = 1; i < KG_OI_BUFSZ; i++) {
        before |= e[i - 1];
        le[i] = found ? !before : i <= u.n;
    }
    int tc[KG_OI_BUFSZ], to[KG_OI_BUFSZ]; // what slot i holds if the new entry lands in front of it (independent of C)
#pragma unroll
    for (int i = 1; i < KG_OI_BUFSZ; i++) {
        tc[i] = le[i] ? u.c[i - 1] : u.c[i];
        to[i] = le[i] ? u.o[i - 1] : u.o[i];
    }
    bool g[KG_OI_BUFSZ];
#pragma unroll
    for (int i = 0; i < KG_OI_BUFSZ - 1; i++) g[i] = u.c[i] > C;
    g[KG_OI_BUFSZ - 1] = false;
    int nc[KG_OI_BUFSZ], no[KG_OI_BUFSZ];
    nc[0] = g[0] ? u.c[0] : C;
    no[0] = g[0] ? u.o[0] : oI;
#pragma unroll
    for (int i = 1; i < KG_OI_BUFSZ; i++) {
        nc[i] = g[i] ? u.c[i] : (g[i - 1] ? C : tc[i]);
        no[i] = g[i] ? u.o[i] : (g[i - 1] ? oI : to[i]);
    }
#pragma unroll
    for (int i = 0; i < KG_OI_BUFSZ; i++) {
        u.c[i] = nc[i];
        u.o[i] = no[i];
    }
    u.n += (!found && u.n < KG_OI_BUFSZ);
}

struct KgFsmParams {
    int min_hits, max_gap, order_constraint;
    float min_weighted;
};

struct KgHitLite {
    int pos, fI, avg, oI;
    float wt;
};

struct KgDevCall { // one CALL before it is tagged with (seq, strand_frame)
    int start, end, count, fI;
    float weighted;
    int hits_before;
};

struct KgFsm {
    // open run
    int n, cur, first_pos, cnt, last_match;
    float w;
    KgHitLite l1, l2;
    KgOtuBuf otu_c, otu_t;
    int pend_oI, pend_cnt;
    int consumed;       // HIT lines so far in this container
    int ncalls;

    __device__ __forceinline__ void begin_sequence() { kg_otu_clear(otu_c); }
    __device__ __forceinline__ void begin_container() {
        n = 0;
        cur = 0; // KGJ:467
        consumed = 0;
        ncalls = 0;
        cnt = 0;
        w = 0.f;
        first_pos = last_match = 0;
        l1 = KgHitLite{0, 0, 0, 0, 0.f};
        l2 = l1;
        pend_oI = pend_cnt = 0;
    }
    __device__ __forceinline__ void count_otu(int oI) {
        if (pend_cnt && oI == pend_oI) {
            pend_cnt++;
        } else {
            if (pend_cnt) kg_otu_update_n(otu_t, pend_oI, pend_cnt);
            pend_oI = oI;
            pend_cnt = 1;
        }
    }

    __device__ __forceinline__ void append(const KgHitLite& h) {
        if (n == 0) {
            first_pos = h.pos;
            cnt = 0;
            w = 0.f;
            otu_t = otu_c;
            pend_cnt = 0;
        }
        n++;
        l2 = l1;
        l1 = h;
        if (h.fI == cur) { // KGJ:391-395
            cnt++;
            w = __fadd_rn(w, h.wt);
            last_match = h.pos;
            count_otu(h.oI);
        }
    }

    // processSetOfHits, KGJ:385-455
    template <class Emit>
    __device__ __forceinline__ void process(const KgFsmParams& p, Emit& emit) {
        if (cnt >= p.min_hits && w >= p.min_weighted) { // KGJ:397
            KgDevCall c = {first_pos, last_match + (KG_K - 1), cnt, cur, w, consumed};
            emit(ncalls, c);
            ncalls++;
            if (pend_cnt) kg_otu_update_n(otu_t, pend_oI, pend_cnt);
            otu_c = otu_t; // the replay of KGJ:413-439 has already been done incrementally
        }
        pend_cnt = 0;
        if (n >= 2 && l2.fI != cur && l2.fI == l1.fI) { // KGJ:442-449: the pair seeds the next run
            cur = l1.fI;
            n = 2;
            first_pos = l2.pos;
            cnt = 2;
            w = __fadd_rn(__fadd_rn(0.f, l2.wt), l1.wt);
            last_match = l1.pos;
            otu_t = otu_c;
            count_otu(l2.oI);
            count_otu(l1.oI);
        } else {
            n = 0; // KGJ:452
        }
    }

    // body of the for loop of gatherHits, KGJ:468-510
    template <class Emit>
    __device__ __forceinline__ void hit(const KgFsmParams& p, const KgHitLite& h, Emit& emit) {
        // Fast path -- the hit extends the open run of its own function (most hits inside a gene): no gap (KGJ:477), the
        // list is not empty so currentFI stays (KGJ:486), no order constraint (KGJ:490), room in the list (KGJ:496), and
        // the pair-switch test cannot fire because currentFI == ph.fI (KGJ:503).  What is left of the general path below
        // is exactly these statements; one predictable branch instead of eight.
        if (n > 0 && n < KG_MAX_HITS_PER_SEQ - 2 && h.fI == cur && !p.order_constraint &&
            (int)((unsigned)l1.pos + (unsigned)p.max_gap) >= h.pos) {
            consumed++;
            n++;
            l2 = l1;
            l1 = h;
            cnt++;
            w = __fadd_rn(w, h.wt);
            last_match = h.pos;
            count_otu(h.oI);
            return;
        }
        consumed++;
        if (n > 0 && (int)((unsigned)l1.pos + (unsigned)p.max_gap) < h.pos) { // KGJ:477-484 (Java int wrap-around kept)
            if (n >= p.min_hits) process(p, emit);
            else n = 0;
        }
        if (n == 0) cur = h.fI; // KGJ:486-488
        bool accept = !p.order_constraint || n == 0;
        if (!accept) { // KGJ:491-494
            int d = (int)((unsigned)(h.pos - l1.pos) - (unsigned)(l1.avg - h.avg));
            int ad = d < 0 ? (int)(0u - (unsigned)d) : d;
            accept = (h.fI == l1.fI) && ad <= 20;
        }
        if (accept) {
            if (n < KG_MAX_HITS_PER_SEQ - 2) append(h); // KGJ:496-497
            if (n > 1 && cur != h.fI && l2.fI == l1.fI) process(p, emit); // KGJ:503-507 (tested even when not appended)
        }
    }

    template <class Emit>
    __device__ __forceinline__ void end_container(const KgFsmParams& p, Emit& emit) {
        if (n >= p.min_hits) process(p, emit); // KGJ:511-513
    }
};

// ---------------------------------------------------------------------------------------------------------------
// Segment variant (long contigs).  After a gap larger than max_gap the open run is always empty: at the gap test
// processSetOfHits either clears the list or is entered with a list whose last two entries do NOT form a foreign pair
// (the pair-switch test of KGJ:503-507 has fired on every append, so "l2.fI == l1.fI != cur" never survives a step),
// hence its retained-pair branch (KGJ:442-449) cannot be taken there.  A container's hits can therefore be cut at gaps
// > max_gap into independent segments, one thread each.  The OTU buffer is order-dependent across calls (KGJ:413-438),
// so this FSM only LISTS, per call, the OTU indices of the hits it counts (run-length encoded) and a per-sequence fold
// applies them in order afterwards.
// ---------------------------------------------------------------------------------------------------------------
struct KgSegRuns { // where a segment's OTU runs go: sparse slots starting at the segment's first hit index
    int2* run; // (OTU index, length)
    uint32_t n; // runs of emitted calls
    __device__ __forceinline__ void put(uint32_t at, int oi, uint32_t m) { run[at] = make_int2(oi, (int)m); }
};

struct KgFsmSeg {
    int n, cur, first_pos, cnt, last_match;
    float w;
    KgHitLite l1, l2;
    int consumed, ncalls;
    // The hits KGJ:413-439 would replay into the OTU buffer if the open run is called = its hits with fI == cur, in
    // order.  They are run-length encoded as they arrive (m updates of one index collapse exactly, kg_otu_update_n):
    // finished runs wait in the output slots behind the committed ones, (ro, rm) is the run still growing.  A call commits
    // them; a run that is dropped just forgets them.  A hit is counted by at most one call, so committed + waiting runs
    // never outnumber the hits seen so far and stay inside the segment's own slots.
    uint32_t tent;
    int ro;
    uint32_t rm;

    __device__ __forceinline__ void begin(int consumed0) {
        n = 0;
        cur = 0;
        consumed = consumed0;
        ncalls = 0;
        cnt = 0;
        w = 0.f;
        first_pos = last_match = 0;
        l1 = KgHitLite{0, 0, 0, 0, 0.f};
        l2 = l1;
        tent = 0;
        ro = 0;
        rm = 0;
    }
    __device__ __forceinline__ void count_oi(int oI, KgSegRuns& runs) {
        if (rm && oI == ro) {
            rm++;
        } else {
            if (rm) runs.put(runs.n + tent++, ro, rm);
            ro = oI;
            rm = 1;
        }
    }
    __device__ __forceinline__ void append(const KgHitLite& h, KgSegRuns& runs) {
        if (n == 0) {
            first_pos = h.pos;
            cnt = 0;
            w = 0.f;
            tent = 0;
            rm = 0;
        }
        n++;
        l2 = l1;
        l1 = h;
        if (h.fI == cur) {
            cnt++;
            w = __fadd_rn(w, h.wt);
            last_match = h.pos;
            count_oi(h.oI, runs);
        }
    }
    template <class Emit>
    __device__ __forceinline__ void process(const KgFsmParams& p, Emit& emit, KgSegRuns& runs) {
        if (cnt >= p.min_hits && w >= p.min_weighted) {
            KgDevCall c = {first_pos, last_match + (KG_K - 1), cnt, cur, w, consumed};
            emit(ncalls, c);
            ncalls++;
            if (rm) runs.put(runs.n + tent++, ro, rm);
            runs.n += tent;
        }
        tent = 0;
        rm = 0;
        if (n >= 2 && l2.fI != cur && l2.fI == l1.fI) { // the retained pair seeds the next run (KGJ:442-449)
            cur = l1.fI;
            n = 2;
            first_pos = l2.pos;
            cnt = 2;
            w = __fadd_rn(__fadd_rn(0.f, l2.wt), l1.wt);
            last_match = l1.pos;
            count_oi(l2.oI, runs);
            count_oi(l1.oI, runs);
        } else {
            n = 0;
        }
    }
    template <class Emit>
    __device__ __forceinline__ void hit(const KgFsmParams& p, const KgHitLite& h, Emit& emit, KgSegRuns& runs) {
        // fast path: see KgFsm::hit
        if (n > 0 && n < KG_MAX_HITS_PER_SEQ - 2 && h.fI == cur && !p.order_constraint &&
            (int)((unsigned)l1.pos + (unsigned)p.max_gap) >= h.pos) {
            consumed++;
            n++;
            l2 = l1;
            l1 = h;
            cnt++;
            w = __fadd_rn(w, h.wt);
            last_match = h.pos;
            count_oi(h.oI, runs);
            return;
        }
        consumed++;
        if (n > 0 && (int)((unsigned)l1.pos + (unsigned)p.max_gap) < h.pos) {
            if (n >= p.min_hits) process(p, emit, runs);
            else n = 0;
        }
        if (n == 0) cur = h.fI;
        bool accept = !p.order_constraint || n == 0;
        if (!accept) {
            int d = (int)((unsigned)(h.pos - l1.pos) - (unsigned)(l1.avg - h.avg));
            int ad = d < 0 ? (int)(0u - (unsigned)d) : d;
            accept = (h.fI == l1.fI) && ad <= 20;
        }
        if (accept) {
            if (n < KG_MAX_HITS_PER_SEQ - 2) append(h, runs);
            if (n > 1 && cur != h.fI && l2.fI == l1.fI) process(p, emit, runs);
        }
    }
    template <class Emit>
    __device__ __forceinline__ void end(const KgFsmParams& p, Emit& emit, KgSegRuns& runs) {
        if (n >= p.min_hits) process(p, emit, runs);
    }
};
