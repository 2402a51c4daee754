// tests/fsm_host/fsm_host.cpp -- the PRODUCT's run state machines (kmergutsjava_b200/csrc/kg_fsm.cuh: KgFsm, KgFsmSeg,
// kg_otu_update_n -- the very header the CUDA kernels k_fsm / k_fsm_seg / k_otu_fold are compiled from) built for the HOST,
// so that the CPU test-suite can drive them with thousands of hit lists and compare with the reference's own source
// (tests/test_fsm_host.py).  Test infrastructure: nothing in the product builds or loads this.
//
// Only what the header does not contain is restated here, as the kernels do it (kg_run.cu):
//   k_fsm        begin_sequence; per container: begin_container, hit() for every hit in position order, end_container; the
//                sequence's OTU buffer is otu_c after the last container
//   segment path a container is cut where a hit lies more than max_gap after its predecessor (the gap test of KGJ:477, Java
//                int arithmetic); every segment runs KgFsmSeg on its own, begin(hits of the container before it), one more HIT
//                line counted before end() when the container goes on; the OTU
//                runs of the calls it emitted are folded per sequence afterwards, in order, with kg_otu_update_n (k_otu_fold)
#include <cuda_runtime.h> // __device__ / __forceinline__ as the host compiler sees them, int2 / make_int2

#include <stdint.h>
#include <vector>

static inline float __fadd_rn(float a, float b) { // the device intrinsic: one IEEE-754 binary32 addition, round to nearest
    volatile float r = a + b;
    return r;
}
#include "../../kmergutsjava_b200/csrc/kg_fsm.cuh"

struct HostCall {
    int32_t container, start, end, count, fI;
    float weighted;
    int32_t hits_before;
};
struct Collect {
    std::vector<HostCall>* out;
    int container;
    void operator()(int, const KgDevCall& c) { out->push_back({container, c.start, c.end, c.count, c.fI, c.weighted, c.hits_before}); }
};

// hits of one sequence: container[i] (non-decreasing), pos / fI / avg / oI / wt, sorted by position inside a container.
// path 0 = KgFsm (one thread per sequence), 1 = KgFsmSeg + fold.  Returns the number of calls (at most max_calls are written).
extern "C" int fsm_host_run(int path, int min_hits, int max_gap, int order_constraint, float min_weighted, int ncontainers, int n,
                            const int32_t* container, const int32_t* pos, const int32_t* fI, const int32_t* avg, const int32_t* oI,
                            const float* wt, HostCall* calls, int max_calls, int32_t* otu_n, int32_t* otu_count, int32_t* otu_oI) {
    const KgFsmParams p = {min_hits, max_gap, order_constraint, min_weighted};
    std::vector<HostCall> out;
    KgOtuBuf result;
    kg_otu_clear(result);
    if (path == 0) {
        KgFsm f;
        f.begin_sequence();
        int i = 0;
        for (int k = 0; k < ncontainers; k++) {
            f.begin_container();
            Collect emit{&out, k};
            for (; i < n && container[i] == k; i++) {
                const KgHitLite h = {pos[i], fI[i], avg[i], oI[i], wt[i]};
                f.hit(p, h, emit);
            }
            f.end_container(p, emit);
        }
        result = f.otu_c;
    } else {
        int i = 0;
        for (int k = 0; k < ncontainers; k++) {
            int first = i;
            while (i < n && container[i] == k) {
                int j = i + 1; // the segment [i, j)
                while (j < n && container[j] == k && !((int)((unsigned)pos[j - 1] + (unsigned)max_gap) < pos[j])) j++;
                std::vector<int2> slots((size_t)(j - i) + 2);
                KgSegRuns runs{slots.data(), 0};
                KgFsmSeg f;
                f.begin(i - first);
                Collect emit{&out, k};
                for (int q = i; q < j; q++) {
                    const KgHitLite h = {pos[q], fI[q], avg[q], oI[q], wt[q]};
                    f.hit(p, h, emit, runs);
                }
                // k_fsm_seg: the run that ends at a gap is processed when the NEXT hit of the container arrives, after that hit's
                // HIT line (KGJ:472-480); only the container's last run is processed after the loop (KGJ:511-513)
                if (j < n && container[j] == k) f.consumed++;
                f.end(p, emit, runs);
                for (uint32_t r = 0; r < runs.n; r++) kg_otu_update_n(result, slots[r].x, slots[r].y); // k_otu_fold
                i = j;
            }
        }
    }
    for (int c = 0; c < (int)out.size() && c < max_calls; c++) calls[c] = out[(size_t)c];
    *otu_n = result.n;
    for (int q = 0; q < KG_OI_BUFSZ; q++) {
        otu_count[q] = result.c[q];
        otu_oI[q] = result.o[q];
    }
    return (int)out.size();
}
