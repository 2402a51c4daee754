/*
 * oracle/kg_oracle.c -- CPU restatement of KmerGutsJava's hot path (plain C).
 *
 * TEST INFRASTRUCTURE ONLY (see kg_oracle.h).  PARITY PIN: the reference
 * holds no golden vectors for this path and no JVM exists in the build image,
 * so the reference's OWN SOURCE is executed instead: tests/java_pin/j2py.py
 * transliterates the unmodified KmerGutsJava.java into Python statement by
 * statement (it knows Java syntax and arithmetic, nothing about k-mers) and
 * the reports KmerGutsJava.main writes that way for the eight configs[0] runs
 * are byte-identical to this oracle's (tests/golden/java_transliteration_pin.json,
 * tests/test_java_transliteration.py, tests/test_oracle_golden.py).  Still
 * never run on a real JVM: tests/java_pin/pin_oracle.sh is the one command
 * for that.  Further pins: the hand-traced KATs in tests/golden/, the
 * independent Python restatement in oracle/kg_oracle_py.py, stream-join ==
 * direct-probe.
 *
 * Every function cites the lines of lib/src/kmergutsjava/KmerGutsJava.java
 * ("KGJ") it restates.  Structure follows the reference's call order:
 * prepareQuery -> addKmers -> comparator sort -> lookup -> gatherHits ->
 * processSetOfHits -> report.
 */
#include "kg_oracle.h"

#include <errno.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <zlib.h>

/* ------------------------------------------------------------------ */
/* small helpers                                                       */
/* ------------------------------------------------------------------ */
static void* xmalloc(size_t n) {
    void* p = malloc(n ? n : 1);
    if (!p) { fprintf(stderr, "kg_oracle: out of memory (%zu bytes)\n", n); abort(); }
    return p;
}
static void* xrealloc(void* q, size_t n) {
    void* p = realloc(q, n ? n : 1);
    if (!p) { fprintf(stderr, "kg_oracle: out of memory (%zu bytes)\n", n); abort(); }
    return p;
}
#define VEC_PUSH(arr, n, cap, val)                                               \
    do {                                                                         \
        if ((n) == (cap)) {                                                      \
            (cap) = (cap) ? (cap) * 2 : 16;                                      \
            (arr) = xrealloc((arr), (cap) * sizeof(*(arr)));                     \
        }                                                                        \
        (arr)[(n)++] = (val);                                                    \
    } while (0)

static void set_err(char* err, size_t errlen, const char* msg, const char* arg) {
    if (err && errlen) snprintf(err, errlen, "%s%s", msg, arg ? arg : "");
}

static int ends_with(const char* s, const char* suf) {
    size_t a = strlen(s), b = strlen(suf);
    return a >= b && memcmp(s + a - b, suf, b) == 0;
}

/* Read a whole (optionally gzip'ed) file.  zlib's gz* layer reads plain files transparently, but the
 * reference keys on the ".gz" suffix (KGJ:347, 764, 927), so do the same. */
static uint8_t* slurp(const char* path, size_t* n_out, char* err, size_t errlen) {
    size_t n = 0, cap = 1 << 20;
    uint8_t* buf = xmalloc(cap);
    if (ends_with(path, ".gz")) {
        gzFile g = gzopen(path, "rb");
        if (!g) { set_err(err, errlen, "cannot open ", path); free(buf); return NULL; }
        gzbuffer(g, 1 << 20);
        for (;;) {
            if (n == cap) { cap *= 2; buf = xrealloc(buf, cap); }
            size_t want = cap - n;
            if (want > (1u << 30)) want = 1u << 30;
            int got = gzread(g, buf + n, (unsigned)want);
            if (got < 0) { set_err(err, errlen, "gzip read error in ", path); gzclose(g); free(buf); return NULL; }
            if (got == 0) break;
            n += (size_t)got;
        }
        gzclose(g);
    } else {
        FILE* f = fopen(path, "rb");
        if (!f) { set_err(err, errlen, "cannot open ", path); free(buf); return NULL; }
        for (;;) {
            if (n == cap) { cap *= 2; buf = xrealloc(buf, cap); }
            size_t got = fread(buf + n, 1, cap - n, f);
            if (got == 0) break;
            n += got;
        }
        fclose(f);
    }
    *n_out = n;
    return buf;
}

/* ------------------------------------------------------------------ */
/* alphabet, complement, codon table                                   */
/* ------------------------------------------------------------------ */

/* KGJ:111-175: the 20 upper-case residue letters in alphabetical order -> 0..19, anything else -> 20. */
int kgo_to_amino_acid_off(int c) {
    switch (c) {
        case 'A': return 0;  case 'C': return 1;  case 'D': return 2;  case 'E': return 3;
        case 'F': return 4;  case 'G': return 5;  case 'H': return 6;  case 'I': return 7;
        case 'K': return 8;  case 'L': return 9;  case 'M': return 10; case 'N': return 11;
        case 'P': return 12; case 'Q': return 13; case 'R': return 14; case 'S': return 15;
        case 'T': return 16; case 'V': return 17; case 'W': return 18; case 'Y': return 19;
        default: return 20;
    }
}

/* KGJ:177-260 (note the lower-case 's' maps to upper-case 'S', KGJ:218-219). */
int kgo_compl(int c) {
    switch (c) {
        case 'a': return 't'; case 'A': return 'T';
        case 'c': return 'g'; case 'C': return 'G';
        case 'g': return 'c'; case 'G': return 'C';
        case 't': case 'u': return 'a';
        case 'T': case 'U': return 'A';
        case 'm': return 'k'; case 'M': return 'K';
        case 'r': return 'y'; case 'R': return 'Y';
        case 'w': return 'w'; case 'W': return 'W';
        case 's': return 'S'; case 'S': return 'S';
        case 'y': return 'r'; case 'Y': return 'R';
        case 'k': return 'm'; case 'K': return 'M';
        case 'b': return 'v'; case 'B': return 'V';
        case 'd': return 'h'; case 'D': return 'H';
        case 'h': return 'd'; case 'H': return 'D';
        case 'v': return 'b'; case 'V': return 'B';
        case 'n': return 'n'; case 'N': return 'N';
        default: return c;
    }
}

/* KGJ:263-272 */
void kgo_rev_comp(const uint8_t* data, size_t n, uint8_t* out) {
    size_t pc = 0;
    for (size_t p = n; p-- > 0;) out[pc++] = (uint8_t)kgo_compl(data[p]);
}

/* KGJ:274-292.  The "bad encoding" branch (KGJ:283-290) is unreachable: eight codes < 20 give < 20^8. */
int64_t kgo_encoded_kmer(const uint8_t* codes, size_t pos) {
    int64_t enc = 0;
    for (int i = 0; i < KGO_K; i++) {
        int add = (int8_t)codes[pos + i]; /* Java byte is signed */
        if (add >= 20) return -1;
        enc = enc * 20 + add;
    }
    return enc;
}

/* KGJ:294-318 */
int kgo_dna_char(int c) {
    switch (c) {
        case 'a': case 'A': return 0;
        case 'c': case 'C': return 1;
        case 'g': case 'G': return 2;
        case 't': case 'u': case 'T': case 'U': return 3;
        default: return 4;
    }
}

/* KGJ:88-93 */
static const char GENETIC_CODE[64] = {
    'K','N','K','N','T','T','T','T','R','S','R','S','I','I','M','I',
    'Q','H','Q','H','P','P','P','P','R','R','R','R','L','L','L','L',
    'E','D','E','D','A','A','A','A','G','G','G','G','V','V','V','V',
    '*','Y','*','Y','S','S','S','S','*','C','W','C','L','F','L','F'};
char kgo_genetic_code(int idx) { return GENETIC_CODE[idx & 63]; }

/* KGJ:320-343.  seq.length - 3 is computed in signed arithmetic there (L < 3 -> no codon). */
void kgo_translate(const uint8_t* seq, size_t L, int off, uint8_t* pseq, uint8_t* piseq, size_t plen) {
    long max = (long)L - 3;
    size_t p = 0;
    for (long i = off; i <= max;) {
        int c1 = kgo_dna_char(seq[i++]);
        int c2 = kgo_dna_char(seq[i++]);
        int c3 = kgo_dna_char(seq[i++]);
        if (c1 < 4 && c2 < 4 && c3 < 4) {
            char prot = GENETIC_CODE[c1 * 16 + c2 * 4 + c3];
            pseq[p] = (uint8_t)prot;
            piseq[p] = (uint8_t)kgo_to_amino_acid_off(prot);
        } else {
            pseq[p] = 'x';
            piseq[p] = 20;
        }
        p++;
    }
    if (p < plen) {
        pseq[p] = 0;
        piseq[p] = 21;
    }
}

/* ------------------------------------------------------------------ */
/* table image                                                         */
/* ------------------------------------------------------------------ */
struct kgo_table {
    uint8_t* bytes;     /* whole decompressed file, header included */
    size_t nbytes;
    int borrowed;
    int64_t num_sigs, entry_size, version; /* KGJ:933-935 */
};

static int64_t le64(const uint8_t* p) { /* KGJ:1107-1126 */
    uint64_t v = 0;
    for (int i = 7; i >= 0; i--) v = (v << 8) | p[i];
    return (int64_t)v;
}
static int32_t le32(const uint8_t* p) { /* KGJ:1097-1105 */
    uint32_t v = (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
    return (int32_t)v;
}
static float lef32(const uint8_t* p) { /* KGJ:1128-1130 */
    int32_t b = le32(p);
    float f;
    memcpy(&f, &b, 4);
    return f;
}

static kgo_table* table_adopt(uint8_t* bytes, size_t nbytes) {
    kgo_table* t = xmalloc(sizeof *t);
    t->bytes = bytes;
    t->nbytes = nbytes;
    t->borrowed = 0;
    t->num_sigs = t->entry_size = t->version = 0;
    if (nbytes >= 24) { /* KGJ:933-935 */
        t->num_sigs = le64(bytes);
        t->entry_size = le64(bytes + 8);
        t->version = le64(bytes + 16);
    }
    return t;
}
kgo_table* kgo_table_open(const char* path, char* err, size_t errlen) {
    size_t n = 0;
    uint8_t* b = slurp(path, &n, err, errlen);
    if (!b) return NULL;
    if (n < 24) { set_err(err, errlen, "EOF in table header: ", path); free(b); return NULL; }
    return table_adopt(b, n);
}
kgo_table* kgo_table_from_memory(const void* bytes, size_t nbytes) {
    uint8_t* b = xmalloc(nbytes);
    memcpy(b, bytes, nbytes);
    return table_adopt(b, nbytes);
}
kgo_table* kgo_table_borrow(const void* bytes, size_t nbytes) {
    kgo_table* t = table_adopt((uint8_t*)bytes, nbytes);
    t->borrowed = 1;
    return t;
}
void kgo_table_free(kgo_table* t) {
    if (!t) return;
    if (!t->borrowed) free(t->bytes);
    free(t);
}
int64_t kgo_table_num_sigs(const kgo_table* t) { return t->num_sigs; }
int64_t kgo_table_entry_size(const kgo_table* t) { return t->entry_size; }
int64_t kgo_table_version(const kgo_table* t) { return t->version; }

/* ------------------------------------------------------------------ */
/* prepareQuery / addKmers                                             */
/* ------------------------------------------------------------------ */
typedef struct { int64_t value; int32_t hit_cnt_id; int32_t prot_pos; } query_kmer; /* KGJ:1200-1204 */

typedef struct { /* KGJ:1213-1219 */
    int32_t oI, pos, avg_off, fI;
    float wt;
} hit_t;

typedef struct { /* KGJ:1262-1266; key = (seq, strand_frame) */
    int32_t seq, strand_frame;
    hit_t* hits;
    size_t n, cap;
} container;

struct kgo_result {
    size_t nseq;
    int per_seq; /* containers per sequence: 1 (aa) or 6 (dna) */
    container* cnt;
    size_t ncnt;
    kgo_hit_rec* hits; size_t nhits;
    kgo_call_rec* calls; size_t ncalls, calls_cap;
    kgo_otu_rec* otus;
    int64_t num_kmers, kmers_found, pos_count;
    int lookup_error;
};

typedef struct { query_kmer* q; size_t n, cap; } qvec;

/* KGJ:900-922.  The loop bound is pIseq.length - K on purpose (aa mode drops the last window, KGJ:912). */
static void add_kmers(const uint8_t* piseq, size_t plen, int32_t hit_cnt_id, qvec* qs) {
    for (long i = 0; i < (long)plen - KGO_K; i++) {
        int64_t v = kgo_encoded_kmer(piseq, (size_t)i);
        if (v < 0) continue;
        query_kmer qk = {v, hit_cnt_id, (int32_t)i};
        VEC_PUSH(qs->q, qs->n, qs->cap, qk);
    }
}

/* KGJ:1051-1074 */
static void prepare_query(const kgo_params* p, int32_t seq_idx, const uint8_t* seq, size_t L, qvec* qs,
                          kgo_result* r) {
    if (p->aa) {
        uint8_t* piseq = xmalloc(L);
        for (size_t i = 0; i < L; i++) piseq[i] = (uint8_t)kgo_to_amino_acid_off(seq[i]);
        container c = {seq_idx, 0, NULL, 0, 0};
        r->cnt[r->ncnt] = c;
        add_kmers(piseq, L, (int32_t)r->ncnt, qs);
        r->ncnt++;
        free(piseq);
    } else {
        size_t len = L / 3 + 1; /* KGJ:1061; buffers reused across the six frames exactly as there */
        uint8_t* pseq = xmalloc(len);
        uint8_t* piseq = xmalloc(len);
        memset(pseq, 0, len);
        memset(piseq, 0, len);
        for (int frame = 0; frame < 3; frame++) {
            kgo_translate(seq, L, frame, pseq, piseq, len);
            container c = {seq_idx, frame, NULL, 0, 0};
            r->cnt[r->ncnt] = c;
            add_kmers(piseq, len, (int32_t)r->ncnt, qs);
            r->ncnt++;
        }
        uint8_t* rc = xmalloc(L);
        kgo_rev_comp(seq, L, rc);
        for (int frame = 0; frame < 3; frame++) {
            kgo_translate(rc, L, frame, pseq, piseq, len);
            container c = {seq_idx, 3 + frame, NULL, 0, 0};
            r->cnt[r->ncnt] = c;
            add_kmers(piseq, len, (int32_t)r->ncnt, qs);
            r->ncnt++;
        }
        free(rc);
        free(pseq);
        free(piseq);
    }
}

/* ------------------------------------------------------------------ */
/* comparator sort (KGJ:1076-1095).  Collections.sort is a stable     */
/* merge sort, so equal keys keep insertion order.                     */
/* ------------------------------------------------------------------ */
static int qk_less_eq(const query_kmer* a, const query_kmer* b, int64_t num_sigs) {
    int64_t h1 = a->value % num_sigs, h2 = b->value % num_sigs; /* KGJ:1086-1087 */
    if (h1 != h2) return h1 < h2;
    return a->value <= b->value; /* KGJ:1090; <= keeps the merge stable */
}
static void qk_merge_sort(query_kmer* a, query_kmer* tmp, size_t n, int64_t num_sigs) {
    if (n < 2) return;
    size_t h = n / 2;
    qk_merge_sort(a, tmp, h, num_sigs);
    qk_merge_sort(a + h, tmp, n - h, num_sigs);
    size_t i = 0, j = h, k = 0;
    while (i < h && j < n) tmp[k++] = qk_less_eq(&a[i], &a[j], num_sigs) ? a[i++] : a[j++];
    while (i < h) tmp[k++] = a[i++];
    while (j < n) tmp[k++] = a[j++];
    memcpy(a, tmp, n * sizeof *a);
}

static void container_add(container* c, hit_t h) { VEC_PUSH(c->hits, c->n, c->cap, h); }

/* ------------------------------------------------------------------ */
/* lookup, faithful variant: KGJ:944-1034                              */
/* ------------------------------------------------------------------ */
typedef struct { int64_t value; size_t first, count; } inprog_t; /* HashMap<Long, List<QueryKmer>> KGJ:962 */

static void lookup_stream_join(const kgo_table* t, const query_kmer* q, size_t nq, kgo_result* r) {
    const int64_t num_sigs = t->num_sigs, entry_size = t->entry_size;
    size_t stream = 24; /* the stream is positioned just after the header (KGJ:933-935) */
    int64_t cur_hash = 0;
    size_t qi = 0; /* kmerStorage.loadNext() cursor (KGJ:865-870) */
    inprog_t* ip = NULL;
    size_t nip = 0, capip = 0;
    /* queries with equal value are adjacent after the sort, so a List<QueryKmer> is a range [first, first+count) */
    while (qi < nq || nip > 0) { /* KGJ:964 */
        int64_t needed = cur_hash;
        if (nip == 0) { /* KGJ:966-974 */
            needed = q[qi].value % num_sigs;
            inprog_t e = {q[qi].value, qi, 1};
            VEC_PUSH(ip, nip, capip, e);
            qi++;
        }
        while (qi < nq) { /* KGJ:976-989 */
            if (q[qi].value % num_sigs != needed) break;
            size_t j;
            for (j = 0; j < nip; j++)
                if (ip[j].value == q[qi].value) break;
            if (j < nip) {
                ip[j].count++;
            } else {
                inprog_t e = {q[qi].value, qi, 1};
                VEC_PUSH(ip, nip, capip, e);
            }
            qi++;
        }
        if (needed > cur_hash) { /* KGJ:991-994: skip() may run past EOF; the next read then fails */
            stream += (size_t)(entry_size * (needed - cur_hash));
            cur_hash = needed;
        }
        if (stream > t->nbytes || t->nbytes - stream < 24) { /* EOFException, KGJ:1102-1103, caught at KGJ:799-802 */
            r->lookup_error = 1;
            break;
        }
        const uint8_t* e = t->bytes + stream; /* KGJ:995-999: always 24 bytes, whatever entrySize says */
        int64_t which_kmer = le64(e);
        int32_t otu_index = le32(e + 8);
        int32_t avg_from_end = le32(e + 12);
        int32_t function_index = le32(e + 16);
        float function_wt = lef32(e + 20);
        stream += 24;
        if (which_kmer > KGO_MAX_ENCODED) { /* KGJ:1000-1001 */
            nip = 0;
        } else {
            for (size_t j = 0; j < nip; j++) { /* KGJ:1004-1016 */
                if (ip[j].value != which_kmer) continue;
                r->kmers_found++;
                for (size_t k = ip[j].first; k < ip[j].first + ip[j].count; k++) {
                    hit_t h = {otu_index, q[k].prot_pos, avg_from_end, function_index, function_wt};
                    container_add(&r->cnt[q[k].hit_cnt_id], h);
                    r->pos_count++;
                }
                ip[j] = ip[nip - 1];
                nip--;
                break;
            }
        }
        cur_hash++; /* KGJ:1018 */
    }
    free(ip);
}

/* ------------------------------------------------------------------ */
/* lookup, direct variant: linear probing WITHOUT wrap into the same   */
/* 24-byte-entry image; running off the end is a miss for that query   */
/* (in the stream join it aborts the pass, by which time every query   */
/* still pending is one that would have run off the end too).          */
/* ------------------------------------------------------------------ */
static void lookup_direct(const kgo_table* t, const query_kmer* q, size_t nq, kgo_result* r) {
    const int64_t num_sigs = t->num_sigs, entry_size = t->entry_size;
    for (size_t i = 0; i < nq; i++) {
        int64_t slot = q[i].value % num_sigs;
        /* slot s sits at byte 24 + entry_size*h + 24*(s-h) in stream terms; with entry_size == 24 that is 24+24*s */
        size_t stream = 24 + (size_t)(entry_size * slot);
        for (;;) {
            if (stream > t->nbytes || t->nbytes - stream < 24) { r->lookup_error = 1; break; }
            const uint8_t* e = t->bytes + stream;
            int64_t which_kmer = le64(e);
            if (which_kmer > KGO_MAX_ENCODED) break;
            if (which_kmer == q[i].value) {
                hit_t h = {le32(e + 8), q[i].prot_pos, le32(e + 12), le32(e + 16), lef32(e + 20)};
                container_add(&r->cnt[q[i].hit_cnt_id], h);
                r->pos_count++;
                break;
            }
            stream += 24;
        }
    }
}

/* ------------------------------------------------------------------ */
/* Java float formatting                                               */
/* ------------------------------------------------------------------ */
/* String.format("%f", Float) widens to double, takes the shortest decimal string that round-trips
 * (FloatingDecimal) and rounds THAT to `prec` places HALF_UP (Formatter / FormattedFloatingDecimal).  C's printf
 * rounds the exact binary value half-to-even, which differs on ties such as 1/128 = 0.0078125. */
void kgo_java_format_f(float v, int prec, char* out, size_t outlen) {
    double d = (double)v;
    if (isnan(d)) { snprintf(out, outlen, "NaN"); return; }
    if (isinf(d)) { snprintf(out, outlen, d < 0 ? "-Infinity" : "Infinity"); return; }
    int neg = signbit(d) != 0;
    d = fabs(d);
    char sci[64];
    int p;
    for (p = 0; p <= 16; p++) { /* shortest round-trip mantissa */
        snprintf(sci, sizeof sci, "%.*e", p, d);
        if (strtod(sci, NULL) == d) break;
    }
    /* sci = D.DDDDe[+-]XX -> digit string + decimal exponent */
    char digits[40];
    int nd = 0;
    char* e = strchr(sci, 'e');
    for (char* c = sci; c < e; c++)
        if (*c >= '0' && *c <= '9') digits[nd++] = *c;
    int exp10 = atoi(e + 1); /* value = 0.d1d2... * 10^(exp10+1) */
    int point = exp10 + 1;   /* number of digits before the decimal point (may be <= 0) */
    /* fixed-point digit buffer: integer part then fraction, long enough for prec+1 fraction digits */
    char fix[400];
    int nf = 0, int_len = point > 0 ? point : 1;
    if (point <= 0) {
        fix[nf++] = '0';
        for (int i = 0; i < -point; i++) fix[nf++] = '0';
        for (int i = 0; i < nd; i++) fix[nf++] = digits[i];
    } else {
        for (int i = 0; i < point; i++) fix[nf++] = i < nd ? digits[i] : '0';
        for (int i = point; i < nd; i++) fix[nf++] = digits[i];
    }
    while (nf < int_len + prec + 1) fix[nf++] = '0';
    /* HALF_UP at int_len+prec */
    int keep = int_len + prec;
    int carry = fix[keep] >= '5';
    for (int i = keep - 1; i >= 0 && carry; i--) {
        if (fix[i] == '9') fix[i] = '0';
        else { fix[i]++; carry = 0; }
    }
    char res[420];
    int nr = 0;
    if (neg) res[nr++] = '-';
    if (carry) res[nr++] = '1';
    for (int i = 0; i < int_len; i++) res[nr++] = fix[i];
    if (prec > 0) {
        res[nr++] = '.';
        for (int i = 0; i < prec; i++) res[nr++] = fix[int_len + i];
    }
    res[nr] = 0;
    snprintf(out, outlen, "%s", res);
}

/* ------------------------------------------------------------------ */
/* gatherHits / processSetOfHits                                       */
/* ------------------------------------------------------------------ */
typedef struct {
    const kgo_params* p;
    const kgo_functions* fn; /* may be NULL (records only) */
    FILE* pw;                /* may be NULL */
    kgo_otu_rec* otu;
    kgo_call_rec** calls; size_t* ncalls; size_t* calls_cap;
    int32_t seq, strand_frame;
    int32_t hits_printed;
} fsm_ctx;

static void display_hits(const hit_t* hits, size_t n, FILE* pw) { /* KGJ:375-383 */
    fputs("hits: ", pw);
    for (size_t i = 0; i < n; i++) {
        char w[64];
        kgo_java_format_f(hits[i].wt, 6, w, sizeof w);
        fprintf(pw, "%d/%s/%d ", hits[i].pos, w, hits[i].fI);
    }
    fputc('\n', pw);
}

/* KGJ:385-455.  `hits`/`*n` is the ArrayList; returns the new currentFI. */
static int32_t process_set_of_hits(fsm_ctx* c, hit_t* hits, size_t* n, int32_t current_fI) {
    int fI_count = 0;
    float weighted_hits = 0;
    size_t last_hit = 0;
    for (size_t i = 0; i < *n; i++) { /* KGJ:390-396 */
        if (hits[i].fI == current_fI) {
            last_hit = i;
            fI_count++;
            weighted_hits += hits[i].wt;
        }
    }
    if (fI_count >= c->p->min_hits && weighted_hits >= (float)c->p->min_weighted_hits) { /* KGJ:397 */
        kgo_call_rec call = {c->seq, c->strand_frame, hits[0].pos, hits[last_hit].pos + (KGO_K - 1),
                             fI_count, current_fI, weighted_hits, c->hits_printed};
        VEC_PUSH(*c->calls, *c->ncalls, *c->calls_cap, call);
        if (c->pw) { /* KGJ:398-404 */
            char w[64];
            kgo_java_format_f(weighted_hits, 6, w, sizeof w);
            const char* name = "";
            if (c->fn && current_fI >= 0 && (size_t)current_fI < c->fn->n) name = c->fn->name[current_fI];
            fprintf(c->pw, "CALL\t%d\t%d\t%d\t%d\t%s\t%s\n", call.start, call.end, fI_count, current_fI, name, w);
            if (c->p->debug) { /* KGJ:406-409 */
                fputs("after-call: ", c->pw);
                display_hits(hits, *n, c->pw);
            }
        }
        kgo_otu_rec* o = c->otu;
        for (size_t i = 0; i <= last_hit; i++) { /* KGJ:413-439 */
            if (hits[i].fI != current_fI) continue;
            int j;
            for (j = 0; j < o->n && o->oI[j] != hits[i].oI; j++) {}
            if (j == o->n) {
                if (o->n == KGO_OI_BUFSZ) j--; /* overwrite the last entry, KGJ:419-421 */
                else o->n++;
                o->oI[j] = hits[i].oI;
                o->count[j] = 1;
            } else {
                o->count[j]++;
            }
            while (j > 0 && o->count[j - 1] <= o->count[j]) { /* KGJ:432-437 */
                int32_t tc = o->count[j - 1], to = o->oI[j - 1];
                o->count[j - 1] = o->count[j]; o->oI[j - 1] = o->oI[j];
                o->count[j] = tc; o->oI[j] = to;
                j--;
            }
        }
    }
    size_t num = *n;
    /* KGJ:442-443.  num == 1 would index -1 there (ArrayIndexOutOfBounds); the ABI requires min_hits >= 2,
     * which makes num >= 2 on every path that reaches here. */
    if (num >= 2 && hits[num - 2].fI != current_fI && hits[num - 2].fI == hits[num - 1].fI) {
        current_fI = hits[num - 1].fI; /* KGJ:444-449: the pair seeds the next run */
        hits[0] = hits[num - 2];
        hits[1] = hits[num - 1];
        *n = 2;
    } else {
        *n = 0; /* KGJ:452 */
    }
    return current_fI;
}

static int hit_pos_less_eq(const hit_t* a, const hit_t* b) { return a->pos <= b->pos; }
static void hit_merge_sort(hit_t* a, hit_t* tmp, size_t n) { /* stable, like Collections.sort KGJ:460-465 */
    if (n < 2) return;
    size_t h = n / 2;
    hit_merge_sort(a, tmp, h);
    hit_merge_sort(a + h, tmp, n - h);
    size_t i = 0, j = h, k = 0;
    while (i < h && j < n) tmp[k++] = hit_pos_less_eq(&a[i], &a[j]) ? a[i++] : a[j++];
    while (i < h) tmp[k++] = a[i++];
    while (j < n) tmp[k++] = a[j++];
    memcpy(a, tmp, n * sizeof *a);
}

/* KGJ:457-514.  all_hits must already be sorted by pos (the sort is done once, right after lookup). */
static void gather_hits(fsm_ctx* c, const hit_t* all_hits, size_t nall) {
    const kgo_params* p = c->p;
    /* the open run never holds more than min(nall, MAX_HITS_PER_SEQ - 2) hits (KGJ:496) */
    hit_t* hits = xmalloc(((nall < KGO_MAX_HITS_PER_SEQ ? nall : KGO_MAX_HITS_PER_SEQ) + 2) * sizeof *hits);
    size_t n = 0;
    int32_t current_fI = 0;
    c->hits_printed = 0;
    for (size_t a = 0; a < nall; a++) {
        const hit_t* ph = &all_hits[a];
        int32_t avg_off_end = ph->avg_off, fI = ph->fI;
        if (c->pw && p->debug) { /* KGJ:472-475 */
            char w[64];
            kgo_java_format_f(ph->wt, 3, w, sizeof w);
            fprintf(c->pw, "HIT\t%d\t%d\t%d\t%d\t%s\t%d\n", ph->pos, 0, avg_off_end, fI, w, ph->oI);
        }
        c->hits_printed++;
        /* KGJ:477-484; Java int arithmetic wraps, keep that */
        if (n > 0 && (int32_t)((uint32_t)hits[n - 1].pos + (uint32_t)p->max_gap) < ph->pos) {
            if ((long)n >= p->min_hits) current_fI = process_set_of_hits(c, hits, &n, current_fI);
            else n = 0;
        }
        if (n == 0) current_fI = fI; /* KGJ:486-488 */
        int accept = !p->order_constraint || n == 0; /* KGJ:490-494 */
        if (!accept) {
            const hit_t* l = &hits[n - 1];
            int32_t d = (int32_t)((uint32_t)(ph->pos - l->pos) - (uint32_t)(l->avg_off - avg_off_end));
            int32_t ad = d < 0 ? (int32_t)(0u - (uint32_t)d) : d; /* Math.abs(Integer.MIN_VALUE) stays negative */
            accept = fI == l->fI && ad <= 20;
        }
        if (accept) {
            if (n < KGO_MAX_HITS_PER_SEQ - 2) { /* KGJ:496-502 */
                hits[n++] = *ph;
                if (c->pw && p->debug) {
                    fputs("after-hit: ", c->pw);
                    display_hits(hits, n, c->pw);
                }
            }
            if (n > 1 && current_fI != fI && hits[n - 2].fI == hits[n - 1].fI) /* KGJ:503-508 */
                current_fI = process_set_of_hits(c, hits, &n, current_fI);
        }
    }
    if ((long)n >= p->min_hits) process_set_of_hits(c, hits, &n, current_fI); /* KGJ:511-513 */
    free(hits);
}

size_t kgo_gather_hits(const kgo_params* p, kgo_hit_rec* hits, size_t nhits, kgo_otu_rec* otu,
                       kgo_call_rec* calls, size_t max_calls) {
    hit_t* h = xmalloc(nhits * sizeof *h);
    hit_t* tmp = xmalloc(nhits * sizeof *tmp);
    for (size_t i = 0; i < nhits; i++) {
        hit_t x = {hits[i].oI, hits[i].pos, hits[i].avg_off_from_end, hits[i].fI, hits[i].function_wt};
        h[i] = x;
    }
    hit_merge_sort(h, tmp, nhits);
    free(tmp);
    kgo_call_rec* cv = NULL;
    size_t nc = 0, cc = 0;
    fsm_ctx c = {p, NULL, NULL, otu, &cv, &nc, &cc, nhits ? hits[0].seq : 0, nhits ? hits[0].strand_frame : 0, 0};
    gather_hits(&c, h, nhits);
    size_t w = nc < max_calls ? nc : max_calls;
    if (w) memcpy(calls, cv, w * sizeof *cv);
    free(cv);
    free(h);
    return nc;
}

/* ------------------------------------------------------------------ */
/* run(): KGJ:742-820 minus file handling                              */
/* ------------------------------------------------------------------ */
kgo_result* kgo_run(const kgo_table* t, const kgo_params* p, const uint8_t* seq_bytes, const uint64_t* offsets,
                    size_t n, int variant) {
    kgo_result* r = xmalloc(sizeof *r);
    memset(r, 0, sizeof *r);
    r->nseq = n;
    r->per_seq = p->aa ? 1 : 6;
    r->cnt = xmalloc(n * (size_t)r->per_seq * sizeof *r->cnt);
    qvec qs = {NULL, 0, 0};
    for (size_t s = 0; s < n; s++) /* readFasta callback, KGJ:778-784 */
        prepare_query(p, (int32_t)s, seq_bytes + offsets[s], (size_t)(offsets[s + 1] - offsets[s]), &qs, r);
    r->num_kmers = (int64_t)qs.n;

    if (t->num_sigs > 0) {
        if (variant == KGO_LOOKUP_STREAM_JOIN) {
            query_kmer* tmp = xmalloc(qs.n * sizeof *tmp);
            qk_merge_sort(qs.q, tmp, qs.n, t->num_sigs); /* finalizeSorting, KGJ:785, 846-847 */
            free(tmp);
            lookup_stream_join(t, qs.q, qs.n, r);
        } else {
            lookup_direct(t, qs.q, qs.n, r);
        }
    } else if (qs.n) {
        r->lookup_error = 1; /* value % 0 -> ArithmeticException, swallowed at KGJ:799-802 */
    }
    free(qs.q);

    /* grouping: KGJ:805-818 */
    r->otus = xmalloc(n * sizeof *r->otus);
    size_t total = 0;
    for (size_t c = 0; c < r->ncnt; c++) total += r->cnt[c].n;
    r->hits = xmalloc(total * sizeof *r->hits);
    size_t maxn = 0;
    for (size_t c = 0; c < r->ncnt; c++)
        if (r->cnt[c].n > maxn) maxn = r->cnt[c].n;
    hit_t* tmp = xmalloc(maxn * sizeof *tmp);
    for (size_t s = 0; s < n; s++) {
        kgo_otu_rec* otu = &r->otus[s];
        memset(otu, 0, sizeof *otu); /* new ArrayList per sequence, KGJ:528, 540 */
        for (int k = 0; k < r->per_seq; k++) {
            container* c = &r->cnt[s * (size_t)r->per_seq + (size_t)k];
            hit_merge_sort(c->hits, tmp, c->n); /* KGJ:460-465 */
            for (size_t i = 0; i < c->n; i++) {
                kgo_hit_rec h = {c->seq, c->strand_frame, c->hits[i].pos, c->hits[i].oI, c->hits[i].avg_off,
                                 c->hits[i].fI, c->hits[i].wt};
                r->hits[r->nhits++] = h;
            }
            fsm_ctx ctx = {p, NULL, NULL, otu, &r->calls, &r->ncalls, &r->calls_cap, c->seq, c->strand_frame, 0};
            gather_hits(&ctx, c->hits, c->n);
        }
    }
    free(tmp);
    return r;
}

/* ------------------------------------------------------------------ */
/* T independent runs of the reference algorithm on T contiguous       */
/* shards of the sequences (what running T JVMs on T FASTA shards      */
/* would do); records are concatenated in sequence order.  The         */
/* reference itself is single-threaded.                                */
/* ------------------------------------------------------------------ */
#include <pthread.h>
typedef struct {
    const kgo_table* t; const kgo_params* p; const uint8_t* seq; const uint64_t* off;
    size_t first, n; int variant; kgo_result* r;
} shard_job;
static void* shard_main(void* a) {
    shard_job* j = a;
    uint64_t* off = xmalloc((j->n + 1) * sizeof *off);
    for (size_t i = 0; i <= j->n; i++) off[i] = j->off[j->first + i] - j->off[j->first];
    j->r = kgo_run(j->t, j->p, j->seq + j->off[j->first], off, j->n, j->variant);
    free(off);
    return NULL;
}
kgo_result* kgo_run_parallel(const kgo_table* t, const kgo_params* p, const uint8_t* seq_bytes, const uint64_t* offsets,
                             size_t n, int variant, int threads) {
    if (threads < 1) threads = 1;
    if ((size_t)threads > n) threads = n ? (int)n : 1;
    shard_job* jobs = xmalloc((size_t)threads * sizeof *jobs);
    pthread_t* th = xmalloc((size_t)threads * sizeof *th);
    size_t s = 0;
    for (int k = 0; k < threads; k++) { /* shards balanced by residues */
        uint64_t target = offsets[n] / (uint64_t)threads * (uint64_t)(k + 1);
        size_t e = s;
        if (k == threads - 1) e = n;
        else while (e < n && offsets[e] < target) e++;
        shard_job j = {t, p, seq_bytes, offsets, s, e - s, variant, NULL};
        jobs[k] = j;
        s = e;
        pthread_create(&th[k], NULL, shard_main, &jobs[k]);
    }
    kgo_result* r = xmalloc(sizeof *r);
    memset(r, 0, sizeof *r);
    r->nseq = n;
    r->per_seq = p->aa ? 1 : 6;
    for (int k = 0; k < threads; k++) pthread_join(th[k], NULL);
    size_t nh = 0, nc = 0;
    for (int k = 0; k < threads; k++) { nh += jobs[k].r->nhits; nc += jobs[k].r->ncalls; }
    r->hits = xmalloc(nh * sizeof *r->hits);
    r->calls = xmalloc(nc * sizeof *r->calls);
    r->calls_cap = nc;
    r->otus = xmalloc(n * sizeof *r->otus);
    for (int k = 0; k < threads; k++) {
        kgo_result* q = jobs[k].r;
        for (size_t i = 0; i < q->nhits; i++) { r->hits[r->nhits] = q->hits[i]; r->hits[r->nhits++].seq += (int32_t)jobs[k].first; }
        for (size_t i = 0; i < q->ncalls; i++) { r->calls[r->ncalls] = q->calls[i]; r->calls[r->ncalls++].seq += (int32_t)jobs[k].first; }
        memcpy(r->otus + jobs[k].first, q->otus, q->nseq * sizeof *q->otus);
        r->num_kmers += q->num_kmers; r->kmers_found += q->kmers_found; r->pos_count += q->pos_count;
        r->lookup_error |= q->lookup_error;
        kgo_result_free(q);
    }
    free(jobs);
    free(th);
    return r;
}

void kgo_result_free(kgo_result* r) {
    if (!r) return;
    for (size_t c = 0; c < r->ncnt; c++) free(r->cnt[c].hits);
    free(r->cnt);
    free(r->hits);
    free(r->calls);
    free(r->otus);
    free(r);
}
size_t kgo_result_num_hits(const kgo_result* r) { return r->nhits; }
const kgo_hit_rec* kgo_result_hits(const kgo_result* r) { return r->hits; }
size_t kgo_result_num_calls(const kgo_result* r) { return r->ncalls; }
const kgo_call_rec* kgo_result_calls(const kgo_result* r) { return r->calls; }
size_t kgo_result_num_otus(const kgo_result* r) { return r->nseq; }
const kgo_otu_rec* kgo_result_otus(const kgo_result* r) { return r->otus; }
int64_t kgo_result_num_kmers(const kgo_result* r) { return r->num_kmers; }
int64_t kgo_result_kmers_found(const kgo_result* r) { return r->kmers_found; }
int kgo_result_lookup_error(const kgo_result* r) { return r->lookup_error; }

/* ------------------------------------------------------------------ */
/* text lines: BufferedReader.readLine / String.trim                   */
/* ------------------------------------------------------------------ */
typedef struct { const uint8_t* buf; size_t n, pos; } line_reader;
/* readLine: terminators are \n, \r or \r\n; returns 0 at end of stream */
static int next_line(line_reader* lr, const uint8_t** s, size_t* len) {
    if (lr->pos >= lr->n) return 0;
    size_t a = lr->pos, b = a;
    while (b < lr->n && lr->buf[b] != '\n' && lr->buf[b] != '\r') b++;
    *s = lr->buf + a;
    *len = b - a;
    if (b < lr->n) {
        if (lr->buf[b] == '\r' && b + 1 < lr->n && lr->buf[b + 1] == '\n') b++;
        b++;
    }
    lr->pos = b;
    return 1;
}
/* String.trim(): strips code points <= U+0020 at both ends */
static void trim(const uint8_t** s, size_t* len) {
    while (*len && (*s)[0] <= ' ') { (*s)++; (*len)--; }
    while (*len && (*s)[*len - 1] <= ' ') (*len)--;
}

/* KGJ:345-373 */
kgo_functions* kgo_functions_read(const char* path, char* err, size_t errlen) {
    size_t n = 0;
    uint8_t* buf = slurp(path, &n, err, errlen);
    if (!buf) return NULL;
    kgo_functions* f = xmalloc(sizeof *f);
    f->n = 0;
    f->name = NULL;
    size_t cap = 0;
    line_reader lr = {buf, n, 0};
    const uint8_t* s;
    size_t len;
    for (long line_pos = 0; next_line(&lr, &s, &len); line_pos++) {
        const uint8_t* tab = memchr(s, '\t', len);
        char* end = NULL;
        long idx = -1;
        if (tab) {
            char num[32];
            size_t nl = (size_t)(tab - s) < sizeof num - 1 ? (size_t)(tab - s) : sizeof num - 1;
            memcpy(num, s, nl);
            num[nl] = 0;
            errno = 0;
            idx = strtol(num, &end, 10);
            if (nl == 0 || *end || errno) tab = NULL;
        }
        if (!tab || idx != line_pos) { /* KGJ:361-364 (a missing tab throws a different exception there) */
            char msg[96];
            snprintf(msg, sizeof msg, "Your index must be dense and in order (see line %ld)", line_pos);
            set_err(err, errlen, msg, NULL);
            kgo_functions_free(f);
            free(buf);
            return NULL;
        }
        size_t nl = len - (size_t)(tab + 1 - s);
        char* name = xmalloc(nl + 1);
        memcpy(name, tab + 1, nl);
        name[nl] = 0;
        VEC_PUSH(f->name, f->n, cap, name);
    }
    free(buf);
    return f;
}
void kgo_functions_free(kgo_functions* f) {
    if (!f) return;
    for (size_t i = 0; i < f->n; i++) free(f->name[i]);
    free(f->name);
    free(f);
}

/* KGJ:1132-1192 */
kgo_fasta* kgo_fasta_read(const char* path, char* err, size_t errlen) {
    size_t n = 0;
    uint8_t* buf = slurp(path, &n, err, errlen);
    if (!buf) return NULL;
    kgo_fasta* fa = xmalloc(sizeof *fa);
    fa->n = 0;
    fa->id = NULL;
    fa->seq = xmalloc(n + 1);
    fa->off = NULL;
    size_t idcap = 0, offn = 0, offcap = 0, seqn = 0;
    VEC_PUSH(fa->off, offn, offcap, (uint64_t)0);
    line_reader lr = {buf, n, 0};
    const uint8_t* s1 = NULL;
    size_t l1 = 0;
    int have = 0; /* str1 != null */
    int eof = 0;
    char msg[512];
    for (;;) {
        char* prot_name = NULL;
        if (!have && !eof) { have = next_line(&lr, &s1, &l1); eof = !have; } /* KGJ:1139-1140 */
        for (;;) { /* KGJ:1141-1162 */
            if (!have) break;
            const uint8_t* s2 = s1;
            size_t l2 = l1;
            trim(&s2, &l2);
            if (l2 > 1) {
                const uint8_t* r = s2 + 1;
                size_t rl = l2 - 1;
                trim(&r, &rl);
                if (s2[0] == '>' && rl > 0) {
                    /* StringTokenizer(str2.substring(1), " \t").nextToken() */
                    const uint8_t* a = s2 + 1;
                    const uint8_t* e = s2 + l2;
                    while (a < e && (*a == ' ' || *a == '\t')) a++;
                    const uint8_t* b = a;
                    while (b < e && *b != ' ' && *b != '\t') b++;
                    prot_name = xmalloc((size_t)(b - a) + 1);
                    memcpy(prot_name, a, (size_t)(b - a));
                    prot_name[b - a] = 0;
                    break;
                }
                snprintf(msg, sizeof msg, "Wrong caption line: %.*s", (int)(l2 < 400 ? l2 : 400), s2);
                goto fail;
            }
            have = next_line(&lr, &s1, &l1);
            eof = !have;
        }
        if (!prot_name) break; /* KGJ:1163-1165 */
        for (;;) { /* KGJ:1167-1174 */
            have = next_line(&lr, &s1, &l1);
            eof = !have;
            const uint8_t* s2 = s1;
            size_t l2 = have ? l1 : 0;
            if (have) trim(&s2, &l2);
            if (!have || (l2 > 0 && s2[0] == '>')) {
                snprintf(msg, sizeof msg, "No sequence for caption: %s", prot_name);
                free(prot_name);
                goto fail;
            }
            if (l2 > 0) break;
        }
        for (;;) { /* KGJ:1175-1180: lines are appended UNtrimmed */
            memcpy(fa->seq + seqn, s1, l1);
            seqn += l1;
            have = next_line(&lr, &s1, &l1);
            eof = !have;
            if (!have) break;
            const uint8_t* s2 = s1;
            size_t l2 = l1;
            trim(&s2, &l2);
            if (l2 > 0 && s2[0] == '>') break;
        }
        VEC_PUSH(fa->id, fa->n, idcap, prot_name);
        VEC_PUSH(fa->off, offn, offcap, (uint64_t)seqn);
    }
    free(buf);
    return fa;
fail:
    set_err(err, errlen, msg, NULL);
    free(buf);
    kgo_fasta_free(fa);
    return NULL;
}
void kgo_fasta_free(kgo_fasta* f) {
    if (!f) return;
    for (size_t i = 0; i < f->n; i++) free(f->id[i]);
    free(f->id);
    free(f->seq);
    free(f->off);
    free(f);
}

/* ------------------------------------------------------------------ */
/* report: KGJ:805-818 with processAASeq (526-536), processSeq         */
/* (538-558), tabulateOtuDataForContig (516-524)                       */
/* ------------------------------------------------------------------ */
int kgo_write_report(const kgo_result* r, const kgo_params* p, const kgo_fasta* fa, const kgo_functions* fn,
                     const kgo_table* t, FILE* out) {
    if (fa->n != r->nseq) return -1;
    if (p->debug) { /* KGJ:951-954, 1031-1033 (timer/progress lines are not reproduced) */
        fprintf(out, "Kmer-table info: numSigs=%lld, entrySize=%lld, version=%lld\n", (long long)t->num_sigs,
                (long long)t->entry_size, (long long)t->version);
        if (r->lookup_error) fprintf(out, "Error: null\n");
        else fprintf(out, "Kmers found: %lld (pos-count=%lld)\n", (long long)r->kmers_found, (long long)r->pos_count);
    }
    /* queryIdToLen is a LinkedHashMap (first-insertion order, last value); hitCnts.put keeps the LAST container
     * for a repeated (id, strand, frame): KGJ:772, 782, 805-809. */
    size_t n = fa->n;
    size_t* last = xmalloc(n * sizeof *last);
    char* is_first = xmalloc(n);
    for (size_t i = 0; i < n; i++) { /* O(n^2) only over duplicates: compare via a sorted index instead */
        last[i] = i;
        is_first[i] = 1;
    }
    {
        size_t* idx = xmalloc(n * sizeof *idx);
        for (size_t i = 0; i < n; i++) idx[i] = i;
        /* insertion-stable sort of indices by id (simple merge sort on strings) */
        size_t* tmp = xmalloc(n * sizeof *tmp);
        for (size_t w = 1; w < n; w *= 2) {
            for (size_t lo = 0; lo < n; lo += 2 * w) {
                size_t mid = lo + w < n ? lo + w : n, hi = lo + 2 * w < n ? lo + 2 * w : n;
                size_t i = lo, j = mid, k = lo;
                while (i < mid && j < hi) tmp[k++] = strcmp(fa->id[idx[i]], fa->id[idx[j]]) <= 0 ? idx[i++] : idx[j++];
                while (i < mid) tmp[k++] = idx[i++];
                while (j < hi) tmp[k++] = idx[j++];
            }
            memcpy(idx, tmp, n * sizeof *idx);
        }
        for (size_t a = 0; a < n;) {
            size_t b = a + 1;
            while (b < n && strcmp(fa->id[idx[a]], fa->id[idx[b]]) == 0) b++;
            for (size_t k = a; k < b; k++) {
                last[idx[k]] = idx[b - 1];
                is_first[idx[k]] = (k == a);
            }
            a = b;
        }
        free(tmp);
        free(idx);
    }
    hit_t* tmp = NULL;
    for (size_t i = 0; i < n; i++) {
        if (!is_first[i]) continue;
        size_t s = last[i];
        long seq_len = (long)(fa->off[s + 1] - fa->off[s]);
        kgo_otu_rec otu;
        memset(&otu, 0, sizeof otu);
        kgo_call_rec* cv = NULL;
        size_t nc = 0, cc = 0;
        if (p->aa) fprintf(out, "PROTEIN-ID\t%s\t%ld\n", fa->id[i], seq_len);
        else fprintf(out, "processing %s[%ld]\n", fa->id[i], seq_len);
        for (int k = 0; k < r->per_seq; k++) {
            if (!p->aa)
                fprintf(out, "TRANSLATION\t%s\t%ld\t%c\t%d\n", fa->id[i], seq_len, k < 3 ? '+' : '-', k % 3);
            const container* c = &r->cnt[s * (size_t)r->per_seq + (size_t)k];
            fsm_ctx ctx = {p, fn, out, &otu, &cv, &nc, &cc, c->seq, c->strand_frame, 0};
            gather_hits(&ctx, c->hits, c->n);
        }
        fprintf(out, "OTU-COUNTS\t%s[%ld]", fa->id[i], seq_len);
        for (int j = 0; j < otu.n; j++) fprintf(out, "\t%d-%d", otu.count[j], otu.oI[j]);
        fputc('\n', out);
        free(cv);
    }
    free(tmp);
    free(last);
    free(is_first);
    return 0;
}
