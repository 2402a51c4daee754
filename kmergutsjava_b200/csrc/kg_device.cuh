// kg_device.cuh -- device-side primitives shared by the table builder and the probe kernels (sm_100a).
#pragma once

#include "kg_common.cuh"

struct KgBucket { // the key sector of a bucket line
    uint32_t w[8];
};

// L2 eviction policies: the prefilter must stay resident in L2 (evict_last) while the bucket lines, the residue
// stream and the outputs merely pass through it (evict_first).
__device__ __forceinline__ uint64_t kg_policy_evict_last() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t kg_policy_evict_normal() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t kg_policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}

// One 32-byte sector in one instruction: sm_100a has 256-bit global loads (SASS LDG.E.256).  The table is read-only and
// randomly indexed, so bypass L1 allocation.
__device__ __forceinline__ KgBucket kg_load_sector(const uint4* p) {
    KgBucket r;
    asm volatile("ld.global.nc.L1::no_allocate.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r.w[0]), "=r"(r.w[1]), "=r"(r.w[2]), "=r"(r.w[3]), "=r"(r.w[4]), "=r"(r.w[5]), "=r"(r.w[6]),
                   "=r"(r.w[7])
                 : "l"(p));
    return r;
}
__device__ __forceinline__ KgBucket kg_load_bucket(const uint4* lines, uint32_t b) {
    KgBucket r;
    const uint4* p = lines + (size_t)KG_LINE_UINT4 * b;
    asm volatile("ld.global.nc.L1::no_allocate.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r.w[0]), "=r"(r.w[1]), "=r"(r.w[2]), "=r"(r.w[3]), "=r"(r.w[4]), "=r"(r.w[5]), "=r"(r.w[6]),
                   "=r"(r.w[7])
                 : "l"(p));
    return r;
}
__device__ __forceinline__ KgBucket kg_load_bucket_hint(const uint4* lines, uint32_t b, uint64_t policy) {
    KgBucket r;
    const uint4* p = lines + (size_t)KG_LINE_UINT4 * b;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8], %9;"
                 : "=r"(r.w[0]), "=r"(r.w[1]), "=r"(r.w[2]), "=r"(r.w[3]), "=r"(r.w[4]), "=r"(r.w[5]), "=r"(r.w[6]),
                   "=r"(r.w[7])
                 : "l"(p), "l"(policy));
    return r;
}
// Same, asking L2 to bring in the WHOLE 128-byte line (SASS LDG...LTC128B): the payload sectors are then L2 hits for the load
// that follows a key match.
__device__ __forceinline__ KgBucket kg_load_bucket_line(const uint4* lines, uint32_t b, uint64_t policy) {
    KgBucket r;
    const uint4* p = lines + (size_t)KG_LINE_UINT4 * b;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.L2::128B.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8], %9;"
                 : "=r"(r.w[0]), "=r"(r.w[1]), "=r"(r.w[2]), "=r"(r.w[3]), "=r"(r.w[4]), "=r"(r.w[5]), "=r"(r.w[6]),
                   "=r"(r.w[7])
                 : "l"(p), "l"(policy));
    return r;
}
// payload of slot = bucket*6 + lane: same 128-byte line as the key sector
__device__ __forceinline__ int4 kg_load_payload(const uint4* lines, uint32_t slot) {
    const uint32_t b = slot / KG_BUCKET_KEYS, lane = slot - b * KG_BUCKET_KEYS;
    const uint4* p = lines + (size_t)KG_LINE_UINT4 * b + 2 + lane;
    int4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ unsigned long long kg_load_filter_word(const unsigned long long* filter, uint32_t w, uint64_t policy) {
    unsigned long long v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.u64 %0, [%1], %2;" : "=l"(v) : "l"(filter + w), "l"(policy));
    return v;
}

// Bit i of the result is set when slot i of the bucket holds `key`.
__device__ __forceinline__ uint32_t kg_bucket_match(const KgBucket& bk, uint64_t key) {
    const uint32_t lo = (uint32_t)key;
    const uint32_t x = bk.w[6] ^ ((uint32_t)(key >> 32) * 0x00009249u); // field i == 0  <=>  high bits agree
    uint32_t m = 0;
#pragma unroll
    for (int i = 0; i < KG_BUCKET_KEYS; i++) m |= (uint32_t)((bk.w[i] == lo) & (((x >> (3 * i)) & 7u) == 0u)) << i;
    return m;
}

// Full lookup (used where latency does not matter: verification, the rare overflow continuation).
// Returns the slot (bucket*6 + lane) or 0xFFFFFFFF.
__device__ __forceinline__ uint32_t kg_lookup_from(const KgTableView& t, uint64_t key, uint32_t b) {
    const uint32_t last = t.num_buckets + KG_TAIL_BUCKETS - 1;
    for (;;) {
        KgBucket bk = kg_load_bucket(t.lines, b);
        uint32_t m = kg_bucket_match(bk, key);
        if (m) return b * KG_BUCKET_KEYS + (__ffs(m) - 1);
        if (!(bk.w[6] & KG_W6_FLAG) || b == last) return 0xFFFFFFFFu;
        b++;
    }
}
__device__ __forceinline__ uint32_t kg_lookup(const KgTableView& t, uint64_t key) {
    return kg_lookup_from(t, key, kg_home_bucket(key, t.num_buckets));
}
