// kg_host.cpp -- host-side mirror of KmerGutsJava.run()'s CPU parts: FASTA, function.index, report, main() flags.
// (include/kmerguts_host.h explains why this exists in C++: no JVM in this image.)  No lookup / call logic lives here.
#include "../../include/kmerguts_host.h"

#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>

#include <algorithm>
#include <charconv>
#include <chrono>
#include <string>
#include <string_view>
#include <thread>
#include <type_traits>
#include <unordered_map>
#include <vector>

void kg_set_error(const char* fmt, ...); // kg_table.cu

// vector<T> whose resize() leaves new elements uninitialised (the reader's threads write every byte themselves; a
// value-initialising resize of a few hundred MB is a serial memset in front of them)
template <class T>
struct NoInitAlloc : std::allocator<T> {
    template <class U>
    struct rebind {
        using other = NoInitAlloc<U>;
    };
    NoInitAlloc() = default;
    template <class U>
    NoInitAlloc(const NoInitAlloc<U>&) {}
    template <class U>
    void construct(U* p) noexcept(std::is_nothrow_default_constructible<U>::value) {
        ::new (static_cast<void*>(p)) U;
    }
    template <class U, class... A>
    void construct(U* p, A&&... a) {
        ::new (static_cast<void*>(p)) U(std::forward<A>(a)...);
    }
};
struct kg_fasta {
    std::vector<std::string> ids;
    std::vector<uint8_t, NoInitAlloc<uint8_t>> bytes;
    std::vector<uint64_t, NoInitAlloc<uint64_t>> off{0};
};
struct kg_functions {
    std::vector<std::string> names;
};

namespace {

bool has_suffix(const std::string& s, const char* suf) {
    size_t n = strlen(suf);
    return s.size() >= n && s.compare(s.size() - n, n, suf) == 0;
}
bool file_exists(const std::string& p) {
    struct stat st;
    return stat(p.c_str(), &st) == 0;
}

// whole file, gunzipped when the NAME ends in .gz (the reference decides by name: KGJ:347, 764)
bool read_all(const std::string& path, std::string& out) {
    out.clear();
    if (has_suffix(path, ".gz")) {
        gzFile g = gzopen(path.c_str(), "rb");
        if (!g) return false;
        gzbuffer(g, 1 << 20);
        std::vector<char> buf(4u << 20);
        for (;;) {
            int got = gzread(g, buf.data(), (unsigned)buf.size());
            if (got < 0) {
                gzclose(g);
                return false;
            }
            if (got == 0) break;
            out.append(buf.data(), (size_t)got);
        }
        gzclose(g);
        return true;
    }
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) return false;
    struct stat st;
    size_t have = 0, cap = (fstat(fileno(f), &st) == 0 && st.st_size > 0) ? (size_t)st.st_size + 1 : (4u << 20);
    for (;;) { // straight into the string: no staging buffer, no regrowth when the size is known
        out.resize(cap);
        const size_t got = fread(&out[have], 1, cap - have, f);
        have += got;
        if (have < cap) break;
        cap *= 2;
    }
    out.resize(have);
    fclose(f);
    return true;
}

// BufferedReader.readLine(): a line ends at \n, \r or \r\n; the terminator is not part of the line
struct Lines {
    std::string_view text;
    size_t pos = 0;
    bool next(std::string_view& line) {
        if (pos >= text.size()) return false;
        const char* b = text.data() + pos;
        const size_t left = text.size() - pos;
        const char* nl = (const char*)memchr(b, '\n', left);
        size_t e = nl ? (size_t)(nl - b) : left;            // up to the next \n ...
        if (const char* cr = (const char*)memchr(b, '\r', e)) e = (size_t)(cr - b); // ... or an earlier \r
        line = text.substr(pos, e);
        e += pos;
        if (e < text.size()) e += (text[e] == '\r' && e + 1 < text.size() && text[e + 1] == '\n') ? 2 : 1;
        pos = e;
        return true;
    }
};
// String.trim(): drops chars <= ' ' at both ends
std::string_view jtrim(std::string_view s) {
    while (!s.empty() && (unsigned char)s.front() <= ' ') s.remove_prefix(1);
    while (!s.empty() && (unsigned char)s.back() <= ' ') s.remove_suffix(1);
    return s;
}

} // namespace

namespace {

// The reference's reader (KGJ:1132-1192) over text[begin, end).  `begin` is 0 or the start of a caption line and `end` is
// the start of a caption line or the end of the text, so that the state machine enters every range in the state the
// sequential reader would be in.  Returns false with the reference's message in err.  What happens to a record is the
// sink's business: the reader runs twice over every range, once to COUNT (records, sequence bytes) and once to WRITE
// straight into the final arrays at the places the counts of the earlier ranges give -- the same state machine both times.
struct CountSink {
    size_t nseq = 0, nbytes = 0;
    void line(std::string_view l) { nbytes += l.size(); }
    void record(std::string_view) { nseq++; }
};
struct WriteSink {
    kg_fasta* fa;
    size_t seq, byte; // next record index, next byte position
    void line(std::string_view l) {
        memcpy(fa->bytes.data() + byte, l.data(), l.size());
        byte += l.size();
    }
    void record(std::string_view name) {
        fa->ids[seq].assign(name);
        fa->off[++seq] = byte;
    }
};
template <class Sink>
bool parse_fasta_range(std::string_view whole, size_t begin, size_t end, Sink& sink, std::string& err, int* code) {
    Lines in{whole.substr(0, end), begin};
    std::string_view cur;
    bool have = in.next(cur); // str1
    for (;;) {
        // look for the next caption; lines whose trimmed length is <= 1 are skipped (KGJ:1145, 1161)
        std::string_view name;
        bool got_caption = false;
        while (have) {
            std::string_view t = jtrim(cur);
            if (t.size() > 1) {
                if (t[0] == '>' && !jtrim(t.substr(1)).empty()) {
                    std::string_view r = t.substr(1); // StringTokenizer(" \t").nextToken(), KGJ:1147-1148
                    size_t a = 0;
                    while (a < r.size() && (r[a] == ' ' || r[a] == '\t')) a++;
                    size_t b = a;
                    while (b < r.size() && r[b] != ' ' && r[b] != '\t') b++;
                    name = r.substr(a, b - a);
                    got_caption = true;
                    break;
                }
                err = "Wrong caption line: " + std::string(t.substr(0, std::min<size_t>(t.size(), 400))); // KGJ:1158
                *code = KG_EFORMAT;
                return false;
            }
            have = in.next(cur);
        }
        if (!got_caption) break; // KGJ:1163-1165
        for (;;) {               // first non-blank line after the caption, KGJ:1167-1174
            have = in.next(cur);
            std::string_view t = have ? jtrim(cur) : std::string_view();
            if (!have && end < whole.size()) t = std::string_view(">"); // the next range starts with a caption line
            if ((!have && end >= whole.size()) || (!t.empty() && t[0] == '>')) {
                err = "No sequence for caption: " + std::string(name); // KGJ:1170
                *code = KG_EFORMAT;
                return false;
            }
            if (!t.empty()) break;
        }
        for (;;) { // KGJ:1175-1180: lines are appended as they are, blanks and inner spaces included
            sink.line(cur);
            have = in.next(cur);
            if (!have) break;
            std::string_view t = jtrim(cur);
            if (!t.empty() && t[0] == '>') break;
        }
        sink.record(name);
    }
    return true;
}

// start of the first line at or after `from` whose first character after leading blanks is '>' (a caption line in either
// state of the reader), or text.size()
size_t next_caption_line(std::string_view text, size_t from) {
    size_t p = from;
    if (p == 0) return 0;
    for (;;) {
        // move to the start of the next line: a line ends at \n, \r or \r\n
        const char* b = text.data() + p;
        const size_t left = text.size() - p;
        const char* nl = (const char*)memchr(b, '\n', left);
        size_t e = nl ? (size_t)(nl - b) : left;
        if (const char* cr = (const char*)memchr(b, '\r', e)) e = (size_t)(cr - b);
        p += e;
        if (p >= text.size()) return text.size();
        p += (text[p] == '\r' && p + 1 < text.size() && text[p + 1] == '\n') ? 2 : 1;
        if (p >= text.size()) return text.size();
        size_t q = p;
        while (q < text.size() && (unsigned char)text[q] <= ' ' && text[q] != '\n' && text[q] != '\r') q++;
        if (q < text.size() && text[q] == '>') return p;
    }
}

} // namespace

// ---------------------------------------------------------------------------------------------------------------
// readFasta, KGJ:1132-1192.  The text is cut at caption lines into one range per thread; a caption line puts the
// reference's reader into the same state wherever it comes from, so the ranges parse independently and concatenate.
// ---------------------------------------------------------------------------------------------------------------
static int parse_text(std::string_view text, size_t end, kg_fasta** out) { // text[0, end); a caption line starts at `end` if end < size
    size_t min_chunk = 1u << 20;
    if (const char* e = getenv("KG_FASTA_CHUNK")) min_chunk = std::max<size_t>(1, (size_t)atoll(e)); // tests: many tiny ranges
    unsigned hw = std::thread::hardware_concurrency();
    size_t want = std::min<size_t>({(size_t)(hw ? hw : 1), (size_t)32, end / min_chunk + 1});
    std::vector<size_t> cut{0};
    for (size_t k = 1; k < want; k++) {
        size_t p = next_caption_line(text.substr(0, end), std::max(cut.back() + 1, end / want * k));
        if (p >= end) break;
        if (p > cut.back()) cut.push_back(p);
    }
    cut.push_back(end);
    const size_t parts = cut.size() - 1;
    auto in_parallel = [&](auto&& work) {
        if (parts == 1) {
            work(0);
            return;
        }
        std::vector<std::thread> th;
        for (size_t k = 1; k < parts; k++) th.emplace_back(work, k);
        work(0);
        for (auto& t : th) t.join();
    };
    // pass 1: how many records and sequence bytes every range holds (and whether it parses at all)
    std::vector<CountSink> cnt(parts);
    std::vector<std::string> errs(parts);
    std::vector<int> codes(parts, KG_OK);
    in_parallel([&](size_t k) {
        if (!parse_fasta_range(text, cut[k], cut[k + 1], cnt[k], errs[k], &codes[k]) && codes[k] == KG_OK) codes[k] = KG_EFORMAT;
    });
    for (size_t k = 0; k < parts; k++)
        if (codes[k] != KG_OK) { // the first failing range holds the error the sequential reader would have met first
            kg_set_error("%s", errs[k].c_str());
            return codes[k];
        }
    size_t nseq = 0, nbytes = 0;
    std::vector<size_t> seq0(parts), byte0(parts);
    for (size_t k = 0; k < parts; k++) {
        seq0[k] = nseq;
        byte0[k] = nbytes;
        nseq += cnt[k].nseq;
        nbytes += cnt[k].nbytes;
    }
    // pass 2: every range writes its records where they belong (the pages of the arrays are first touched by the writers)
    kg_fasta* fa = new kg_fasta();
    fa->ids.resize(nseq);
    fa->bytes.resize(nbytes);
    fa->off.resize(nseq + 1);
    fa->off[0] = 0;
    in_parallel([&](size_t k) {
        WriteSink w{fa, seq0[k], byte0[k]};
        parse_fasta_range(text, cut[k], cut[k + 1], w, errs[k], &codes[k]);
    });
    *out = fa;
    return KG_OK;
}

// a plain file is mapped, not copied (the parser only needs a view of the text)
struct MappedText {
    const char* p = nullptr;
    size_t n = 0;
    bool mapped = false;
    std::string owned;
    std::string_view view() const { return mapped ? std::string_view(p, n) : std::string_view(owned); }
    ~MappedText() {
        if (mapped && p) munmap(const_cast<char*>(p), n);
    }
    bool open(const std::string& path) {
        if (!has_suffix(path, ".gz")) {
            const int fd = ::open(path.c_str(), O_RDONLY);
            if (fd < 0) return false;
            struct stat st;
            if (fstat(fd, &st) == 0 && S_ISREG(st.st_mode) && st.st_size > 0) {
                void* m = mmap(nullptr, (size_t)st.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
                if (m != MAP_FAILED) {
                    madvise(m, (size_t)st.st_size, MADV_SEQUENTIAL);
                    p = static_cast<const char*>(m);
                    n = (size_t)st.st_size;
                    mapped = true;
                    ::close(fd);
                    return true;
                }
            }
            ::close(fd);
        }
        return read_all(path, owned); // .gz, pipes, empty files, or no mmap
    }
};

extern "C" int kg_fasta_read(const char* path, kg_fasta** out) {
    if (!path || !out) {
        kg_set_error("kg_fasta_read: null argument");
        return KG_EINVAL;
    }
    MappedText text;
    if (!text.open(path)) {
        kg_set_error("cannot read %s", path);
        return KG_EIO;
    }
    return parse_text(text.view(), text.view().size(), out);
}

// ---------------------------------------------------------------------------------------------------------------
// The same reader over a file that is consumed in batches of whole records: the text is cut at caption
// lines, like the ranges of the parallel reader, so every batch parses as the sequential reader would have parsed it and
// the concatenation of the batches equals kg_fasta_read of the whole file.  Memory is bounded by the batch size (plus
// the longest single record).
// ---------------------------------------------------------------------------------------------------------------
struct kg_fasta_stream {
    gzFile g = nullptr; // gzopen reads plain files too, but the reference decides by NAME (KGJ:764): a plain FILE otherwise
    FILE* f = nullptr;
    std::string buf;
    size_t batch = 0;
    bool eof = false;
    bool fill(size_t upto) {
        std::vector<char> tmp(4u << 20);
        while (!eof && buf.size() < upto) {
            long got = g ? (long)gzread(g, tmp.data(), (unsigned)tmp.size()) : (long)fread(tmp.data(), 1, tmp.size(), f);
            if (got < 0) return false;
            if (got == 0) eof = true;
            else buf.append(tmp.data(), (size_t)got);
        }
        return true;
    }
};
extern "C" int kg_fasta_stream_open(const char* path, size_t batch_bytes, kg_fasta_stream** out) {
    if (!path || !out) {
        kg_set_error("kg_fasta_stream_open: null argument");
        return KG_EINVAL;
    }
    kg_fasta_stream* s = new kg_fasta_stream();
    s->batch = std::max<size_t>(batch_bytes, 1);
    if (has_suffix(path, ".gz")) {
        s->g = gzopen(path, "rb");
        if (s->g) gzbuffer(s->g, 1 << 20);
    } else {
        s->f = fopen(path, "rb");
    }
    if (!s->g && !s->f) {
        delete s;
        kg_set_error("cannot read %s", path);
        return KG_EIO;
    }
    *out = s;
    return KG_OK;
}
extern "C" int kg_fasta_stream_next(kg_fasta_stream* s, kg_fasta** batch) {
    if (!s || !batch) {
        kg_set_error("kg_fasta_stream_next: null argument");
        return KG_EINVAL;
    }
    *batch = nullptr;
    size_t want = s->batch;
    size_t cut;
    for (;;) {
        if (!s->fill(want + 1)) {
            kg_set_error("kg_fasta_stream_next: read error");
            return KG_EIO;
        }
        if (s->buf.empty()) return KG_OK; // end of the file
        if (s->eof && s->buf.size() <= want) {
            cut = s->buf.size();
            break;
        }
        cut = next_caption_line(s->buf, std::min(want, s->buf.size() - 1));
        // The candidate must be followed by at least one more byte of text: a '>' that is the last byte read so far is a
        // caption line whichever bytes follow, but its record has to stay whole in the NEXT batch, which it does (cut before it).
        if (cut < s->buf.size()) break;
        if (s->eof) {
            cut = s->buf.size();
            break;
        }
        want = std::max(want * 2, s->buf.size() + (4u << 20)); // a record longer than the batch: read on to its end
    }
    if (cut == 0) cut = s->buf.size(); // (only when the buffer is one record without a further caption)
    int rc = parse_text(s->buf, cut, batch);
    if (rc != KG_OK) return rc;
    s->buf.erase(0, cut);
    return KG_OK;
}
extern "C" void kg_fasta_stream_close(kg_fasta_stream* s) {
    if (!s) return;
    if (s->g) gzclose(s->g);
    if (s->f) fclose(s->f);
    delete s;
}

extern "C" size_t kg_fasta_count(const kg_fasta* f) { return f ? f->ids.size() : 0; }
extern "C" const char* kg_fasta_id(const kg_fasta* f, size_t i) { return (f && i < f->ids.size()) ? f->ids[i].c_str() : nullptr; }
extern "C" const uint8_t* kg_fasta_bytes(const kg_fasta* f) { return f ? f->bytes.data() : nullptr; }
extern "C" const uint64_t* kg_fasta_offsets(const kg_fasta* f) { return f ? f->off.data() : nullptr; }
extern "C" void kg_fasta_free(kg_fasta* f) { delete f; }

// ---------------------------------------------------------------------------------------------------------------
// loadIndexedArray, KGJ:345-373
// ---------------------------------------------------------------------------------------------------------------
extern "C" int kg_functions_read(const char* path, kg_functions** out) {
    if (!path || !out) {
        kg_set_error("kg_functions_read: null argument");
        return KG_EINVAL;
    }
    std::string text;
    if (!read_all(path, text)) {
        kg_set_error("cannot read %s", path);
        return KG_EIO;
    }
    kg_functions* fn = new kg_functions();
    Lines in{text};
    std::string_view line;
    for (long line_pos = 0; in.next(line); line_pos++) {
        size_t tab = line.find('\t');
        long idx = -1;
        if (tab != std::string_view::npos && tab > 0) {
            auto r = std::from_chars(line.data(), line.data() + tab, idx);
            if (r.ec != std::errc() || r.ptr != line.data() + tab) idx = -1;
        }
        if (idx != line_pos) { // KGJ:361-364
            kg_set_error("Your index must be dense and in order (see line %ld)", line_pos);
            delete fn;
            return KG_EFORMAT;
        }
        fn->names.emplace_back(line.substr(tab + 1));
    }
    *out = fn;
    return KG_OK;
}
extern "C" int kg_functions_load(const char* data_dir, kg_functions** out) {
    if (!data_dir || !out) {
        kg_set_error("kg_functions_load: null argument");
        return KG_EINVAL;
    }
    std::string base = std::string(data_dir) + "/function.index";
    if (file_exists(base + ".gz")) return kg_functions_read((base + ".gz").c_str(), out); // KGJ:754-758
    return kg_functions_read(base.c_str(), out);
}
extern "C" size_t kg_functions_count(const kg_functions* f) { return f ? f->names.size() : 0; }
extern "C" const char* kg_functions_name(const kg_functions* f, size_t i) { return (f && i < f->names.size()) ? f->names[i].c_str() : nullptr; }
extern "C" void kg_functions_free(kg_functions* f) { delete f; }

// ---------------------------------------------------------------------------------------------------------------
// Java's %f.  Formatter widens the float to double, FloatingDecimal yields the shortest digit string that
// round-trips, and that DECIMAL string is rounded HALF_UP to `precision` places -- so 1/128 = 0.0078125 prints as
// 0.007813 where C's printf (exact binary value, ties to even) prints 0.007812.
// ---------------------------------------------------------------------------------------------------------------
static std::string java_f(float v, int prec) {
    double d = (double)v;
    if (std::isnan(d)) return "NaN";
    if (std::isinf(d)) return d < 0 ? "-Infinity" : "Infinity";
    bool neg = std::signbit(d);
    d = std::fabs(d);
    char buf[64];
    auto r = std::to_chars(buf, buf + sizeof buf, d, std::chars_format::scientific); // shortest round-trip
    std::string_view s(buf, (size_t)(r.ptr - buf));
    size_t epos = s.find('e');
    std::string digits;
    for (char ch : s.substr(0, epos))
        if (ch >= '0' && ch <= '9') digits.push_back(ch);
    int exp10 = atoi(std::string(s.substr(epos + 1)).c_str());
    int point = exp10 + 1; // digits before the decimal point
    std::string ip, fp;
    if (point <= 0) {
        ip = "0";
        fp.assign((size_t)(-point), '0');
        fp += digits;
    } else if ((size_t)point >= digits.size()) {
        ip = digits + std::string((size_t)point - digits.size(), '0');
    } else {
        ip = digits.substr(0, (size_t)point);
        fp = digits.substr((size_t)point);
    }
    if (fp.size() < (size_t)prec + 1) fp.append((size_t)prec + 1 - fp.size(), '0');
    bool up = fp[(size_t)prec] >= '5';
    std::string all = ip + fp.substr(0, (size_t)prec);
    for (size_t i = all.size(); up && i-- > 0;) {
        if (all[i] == '9') all[i] = '0';
        else {
            all[i]++;
            up = false;
        }
    }
    if (up) all.insert(all.begin(), '1');
    size_t il = all.size() - (size_t)prec;
    std::string res = neg ? "-" : "";
    res += all.substr(0, il);
    if (prec > 0) {
        res += '.';
        res += all.substr(il);
    }
    return res;
}
extern "C" int kg_format_java_f(float v, int precision, char* out, size_t outlen) {
    if (!out || !outlen || precision < 0 || precision > 30) return KG_EINVAL;
    std::string s = java_f(v, precision);
    snprintf(out, outlen, "%s", s.c_str());
    return KG_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// report, KGJ:810-818
// ---------------------------------------------------------------------------------------------------------------
struct kg_report {
    FILE* out = nullptr;
    bool own = false, header_done = false;
    std::string path;
};
extern "C" int kg_report_open(const char* path, kg_report** out) {
    if (!out) {
        kg_set_error("kg_report_open: null argument");
        return KG_EINVAL;
    }
    kg_report* r = new kg_report();
    r->out = path ? fopen(path, "w") : stdout;
    r->own = path != nullptr;
    r->path = path ? path : "stdout";
    if (!r->out) {
        delete r;
        kg_set_error("cannot write %s", path);
        return KG_EIO;
    }
    *out = r;
    return KG_OK;
}
extern "C" int kg_report_close(kg_report* r) {
    if (!r) return KG_OK;
    int rc = KG_OK;
    if (r->own ? fclose(r->out) != 0 : fflush(r->out) != 0) {
        kg_set_error("kg_report_close: write to %s failed", r->path.c_str());
        rc = KG_EIO;
    }
    delete r;
    return rc;
}
extern "C" int kg_call_dna_range(const kg_call* c, uint64_t contig_len, uint64_t* begin, uint64_t* end, char* strand) {
    if (!c || !begin || !end || !strand) {
        kg_set_error("kg_call_dna_range: null argument");
        return KG_EINVAL;
    }
    if (c->strand_frame < 0 || c->strand_frame > 5 || c->start < 0 || c->end < c->start) {
        kg_set_error("kg_call_dna_range: not a call of a 6-frame run (strand_frame %d, %d..%d)", c->strand_frame, c->start, c->end);
        return KG_EINVAL;
    }
    // residue r of frame f is the codon at nucleotides f + 3r .. f + 3r + 2 of the strand that was translated (KGJ:320-343);
    // the minus frames are frames of the full reverse complement (KGJ:1068-1071): x on it is L - 1 - x on the contig
    const uint64_t f = (uint64_t)(c->strand_frame % 3), a = f + 3ull * (uint64_t)c->start, b = f + 3ull * (uint64_t)c->end + 2;
    if (b >= contig_len) {
        kg_set_error("kg_call_dna_range: residues %d..%d of frame %d lie outside a contig of %llu nucleotides", c->start, c->end,
                     c->strand_frame, (unsigned long long)contig_len);
        return KG_ERANGE;
    }
    const bool minus = c->strand_frame >= 3;
    *strand = minus ? '-' : '+';
    *begin = minus ? contig_len - 1 - b : a;
    *end = minus ? contig_len - 1 - a : b;
    return KG_OK;
}
static int report_add(kg_report* rep, int mode, int flags, const kg_fasta* fa, const kg_functions* fn, const kg_table* table, kg_result* result);
extern "C" int kg_report_add(kg_report* rep, int mode, int flags, const kg_fasta* fa, const kg_functions* fn, const kg_table* table,
                             kg_result* result) {
    if (!rep) {
        kg_set_error("kg_report_add: null argument");
        return KG_EINVAL;
    }
    return report_add(rep, mode, flags, fa, fn, table, result);
}
extern "C" int kg_report_write(const char* path, int mode, int debug, const kg_fasta* fa, const kg_functions* fn,
                               const kg_table* table, kg_result* result) {
    if (!fa || !fn || !result) {
        kg_set_error("kg_report_write: null argument");
        return KG_EINVAL;
    }
    kg_report* rep = nullptr;
    int rc = kg_report_open(path, &rep);
    if (rc != KG_OK) return rc;
    rc = report_add(rep, mode, debug, fa, fn, table, result);
    const int rc2 = kg_report_close(rep);
    return rc != KG_OK ? rc : rc2;
}
static int report_add(kg_report* rep, int mode, int flags, const kg_fasta* fa, const kg_functions* fn, const kg_table* table, kg_result* result) {
    if (!fa || !fn || !result) {
        kg_set_error("kg_report_add: null argument");
        return KG_EINVAL;
    }
    const int debug = flags & 1;
    const bool dna_ranges = (flags & 2) && mode != KG_MODE_AA;
    const char* path = rep->own ? rep->path.c_str() : nullptr;
    const kg_call* calls = nullptr;
    const kg_otu* otus = nullptr;
    const kg_hit* hits = nullptr;
    size_t ncalls = 0, notus = 0, nhits = 0;
    int rc;
    if ((rc = kg_result_calls(result, &calls, &ncalls)) != KG_OK) return rc;
    if ((rc = kg_result_otus(result, &otus, &notus)) != KG_OK) return rc;
    if (debug && (rc = kg_result_hits(result, &hits, &nhits)) != KG_OK) return rc;
    const size_t n = fa->ids.size();
    if (notus != n) {
        kg_set_error("kg_report_write: result has %zu sequences, FASTA has %zu", notus, n);
        return KG_EINVAL;
    }
    // functionArray.get(currentFI) throws for an index outside function.index and aborts the run (KGJ:403): a table that
    // does not belong to this function.index must not turn into CALL lines with an empty name
    for (size_t c = 0; c < ncalls; c++)
        if (calls[c].fI < 0 || (size_t)calls[c].fI >= fn->names.size()) {
            kg_set_error("CALL with function index %d but function.index has %zu names (the reference throws at KGJ:403): table and "
                         "function.index do not belong together", calls[c].fI, fn->names.size());
            return KG_EFORMAT;
        }
    FILE* out = rep->out;
    if (debug && table && !rep->header_done) { // KGJ:951-954
        rep->header_done = true;
        kg_table_info ti;
        kg_table_get_info(table, &ti);
        fprintf(out, "Kmer-table info: numSigs=%lld, entrySize=%lld, version=%lld\n", (long long)ti.num_slots,
                (long long)ti.entry_size, (long long)ti.version);
    }
    // per-sequence ranges into the (seq, strand_frame, pos)-ordered call and hit arrays
    std::vector<size_t> call_lo(n + 1, 0), hit_lo(n + 1, 0);
    {
        size_t c = 0, h = 0;
        for (size_t s = 0; s <= n; s++) {
            while (c < ncalls && calls[c].seq < s) c++;
            while (h < nhits && hits[h].seq < s) h++;
            call_lo[s] = c;
            hit_lo[s] = h;
        }
    }
    // queryIdToLen is a LinkedHashMap<String,Integer>: iteration in FIRST-insertion order with the LAST length;
    // hitCnts.put() keeps the LAST container of a repeated (id, strand, frame).  KGJ:772, 782, 805-809.
    struct FirstLast {
        size_t first, last;
    };
    std::unordered_map<std::string_view, FirstLast> occ;
    occ.reserve(n * 2);
    for (size_t i = 0; i < n; i++) {
        auto r = occ.emplace(std::string_view(fa->ids[i]), FirstLast{i, i});
        if (!r.second) r.first->second.last = i;
    }
    const int per_seq = mode == KG_MODE_AA ? 1 : 6;
    // Sequences are formatted in waves: every thread turns a contiguous block of sequences into text in its own buffer,
    // then the buffers are written in order (fprintf per line was the slowest part of a run; one thread formats ~6 M lines/s).
    auto format_range = [&](size_t i0, size_t i1, std::string& buf) {
        auto put = [&](std::string_view t) { buf.append(t.data(), t.size()); };
        auto num = [&](long long v) {
            char tmp[24];
            auto r = std::to_chars(tmp, tmp + sizeof tmp, v);
            buf.append(tmp, (size_t)(r.ptr - tmp));
        };
        char wbuf[96];
        for (size_t i = i0; i < i1; i++) {
            const std::string& ids = fa->ids[i];
            const FirstLast& fl = occ.find(std::string_view(ids))->second;
            if (fl.first != i) continue;
            const size_t s = fl.last;
            const long long len = (long long)(fa->off[s + 1] - fa->off[s]);
            if (mode == KG_MODE_AA) { // KGJ:529
                put("PROTEIN-ID\t"); put(ids); put("\t"); num(len); put("\n");
            } else { // KGJ:541
                put("processing "); put(ids); put("["); num(len); put("]\n");
            }
            size_t c = call_lo[s], h = hit_lo[s];
            for (int k = 0; k < per_seq; k++) {
                if (mode != KG_MODE_AA) { // KGJ:545-548
                    put("TRANSLATION\t"); put(ids); put("\t"); num(len); put(k < 3 ? "\t+\t" : "\t-\t"); num(k % 3); put("\n");
                }
                int printed = 0;
                auto flush_hits = [&](int upto) { // HIT lines precede the CALL they trigger (KGJ:472-475 before 477-508)
                    while (debug && h < hit_lo[s + 1] && hits[h].strand_frame == k && printed < upto) {
                        kg_format_java_f(hits[h].function_wt, 3, wbuf, sizeof wbuf);
                        put("HIT\t"); num(hits[h].pos); put("\t0\t"); num(hits[h].avg_off_from_end); put("\t"); num(hits[h].fI); put("\t");
                        put(wbuf); put("\t"); num(hits[h].oI); put("\n");
                        h++;
                        printed++;
                    }
                };
                for (; c < call_lo[s + 1] && calls[c].strand_frame == k; c++) {
                    flush_hits(calls[c].hits_before);
                    const kg_call& cl = calls[c];
                    kg_format_java_f(cl.weighted, 6, wbuf, sizeof wbuf);
                    put("CALL\t"); num(cl.start); put("\t"); num(cl.end); put("\t"); num(cl.count); put("\t"); num(cl.fI); put("\t"); // KGJ:398-404
                    put(fn->names[(size_t)cl.fI]);
                    put("\t"); put(wbuf); put("\n");
                    if (dna_ranges) { // extension (-c): the reference prints protein coordinates of the frame only (KGJ:398-401)
                        uint64_t b0 = 0, e0 = 0;
                        char sd = '+';
                        if (kg_call_dna_range(&cl, (uint64_t)len, &b0, &e0, &sd) == KG_OK) {
                            put("DNA-RANGE\t"); num((long long)b0); put("\t"); num((long long)e0); put("\t"); buf.push_back(sd); put("\n");
                        }
                    }
                }
                flush_hits(0x7FFFFFFF);
            }
            put("OTU-COUNTS\t"); put(ids); put("["); num(len); put("]"); // KGJ:518-522
            for (int j = 0; j < otus[s].n; j++) {
                put("\t"); num(otus[s].count[j]); put("-"); num(otus[s].oI[j]);
            }
            put("\n");
        }
    };
    unsigned hw = std::thread::hardware_concurrency();
    const size_t nthreads = std::max<size_t>(1, std::min<size_t>({(size_t)(hw ? hw : 1), (size_t)16, n / 4096 + 1}));
    // a block = about 1/8 of what a thread gets in total, bounded so that a wave's text stays in the tens of megabytes
    const size_t block_cap = mode == KG_MODE_AA ? 65536 : 16; // a contig can carry 10^5 lines
    const size_t block = std::max<size_t>(1, std::min<size_t>((n + nthreads * 8 - 1) / (nthreads * 8), block_cap));
    std::vector<std::string> bufs(nthreads);
    bool io_ok = true;
    for (size_t w0 = 0; w0 < n && io_ok; w0 += block * nthreads) {
        auto work = [&](size_t t) {
            bufs[t].clear();
            const size_t i0 = std::min(n, w0 + t * block), i1 = std::min(n, i0 + block);
            if (i0 < i1) format_range(i0, i1, bufs[t]);
        };
        if (nthreads == 1) {
            work(0);
        } else {
            std::vector<std::thread> th;
            for (size_t t = 1; t < nthreads; t++) th.emplace_back(work, t);
            work(0);
            for (auto& t : th) t.join();
        }
        for (size_t t = 0; t < nthreads && io_ok; t++)
            if (!bufs[t].empty()) io_ok = fwrite(bufs[t].data(), 1, bufs[t].size(), out) == bufs[t].size();
    }
    if (!io_ok) {
        kg_set_error("kg_report_write: write to %s failed", path ? path : "stdout");
        return KG_EIO;
    }
    return KG_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// main, KGJ:560-654
// ---------------------------------------------------------------------------------------------------------------
static void usage() { // KGJ:618-635
    puts("Usage: kmer_guts [options] -D DataDir");
    puts("Arguments:");
    puts(" -a - (optional) amino acids in input FASTA (default is DNA)");
    puts(" -d - (optional) print debug messages");
    puts(" -m - (optional) min. number of hits in result (integer, default = 5)");
    puts(" -M - (optional) min. sum of hit weights (integer, default = 0)");
    puts(" -O - (optional) order constraint (don't use order by default)");
    puts(" -g - (optional) max. gap between hits to be joined (integer, default = 200)");
    puts(" -D - (required) data directory with kmer-table and function-index files");
    puts(" -q - (optional) query fasta file (STDIN if not defined)");
    puts(" -o - (optional) output file (STDOUT if not defined)");
    puts(" -t - (optional) temporary directory (system one is used by default)");
    puts(" -l - (optional) limit for input Kmer array (long, default = 20,000,000)");
    puts(" -G - (extension) CUDA device index (default 0)");
    puts(" -C - (extension) cache file of the GPU table layout: read if present and valid, else written after loading -D");
    puts(" -B - (extension) stream the query in batches of about this many megabytes of FASTA text (bounded memory; an id");
    puts("      repeated in two different batches is reported in both, where the reference keeps only the last)");
    puts(" -c - (extension) after every CALL of a DNA run, a line DNA-RANGE<TAB>begin<TAB>end<TAB>strand (0-based, inclusive,");
    puts("      on the contig as given)");
}

extern "C" int kg_main(int argc, char** argv) {
    kg_params prm;
    kg_params_default(&prm);
    bool aa = false, debug = false;
    const char *dir = nullptr, *query = nullptr, *outp = nullptr, *cache = nullptr;
    int device = 0;
    long batch_mb = 0;
    bool dna_ranges = false;
    std::string err;
    for (int i = 1; i < argc && err.empty(); i++) {
        std::string a = argv[i];
        if (a.empty() || a[0] != '-') { err = "Parameter name should start from '-': " + a; break; } // KGJ:569-572
        if (a.size() != 2) { err = "Unknown parameter: " + a; break; }                               // KGJ:574-576
        auto value = [&]() -> const char* {
            if (i + 1 >= argc) { err = "Missing value for " + a; return "0"; }
            return argv[++i];
        };
        switch (a[1]) {
            case 'a': aa = true; break;
            case 'd': debug = true; break;
            case 'm': prm.min_hits = atoi(value()); break;
            case 'M': prm.min_weighted_hits = atoi(value()); break;
            case 'O': prm.order_constraint = 1; break;
            case 'g': prm.max_gap = atoi(value()); break;
            case 'D': dir = value(); break;
            case 'q': query = value(); break;
            case 'o': outp = value(); break;
            // The reference's -t and -l fall through into `default` and always throw (KGJ:605-610); nothing spills to
            // disk here, so they are accepted and ignored.
            case 't': case 'l': value(); break;
            case 'G': device = atoi(value()); break;
            case 'C': cache = value(); break;
            case 'B': batch_mb = atol(value()); break;
            case 'c': dna_ranges = true; break;
            default: err = "Unknown parameter: " + a;
        }
    }
    if (err.empty() && !dir) err = "-D parameter is required"; // KGJ:613-615
    if (err.empty() && !query) err = "-q parameter is required (reading STDIN is unreachable in the reference too, KGJ:647)";
    if (!err.empty()) {
        // The reference prints this and then carries on into a NullPointerException (KGJ:616-647); exit instead.
        printf("Error: %s\n", err.c_str());
        usage();
        return 2;
    }
    prm.emit_hits = debug ? 1 : 0;
    auto info = [&](const std::string& msg) { // printInfoLine, KGJ:891-898 (the pw part is omitted: wall-clock lines)
        if (outp) puts(msg.c_str());
    };
    auto now = [] { return std::chrono::steady_clock::now(); };
    auto ms = [](auto a, auto b) { return (long long)std::chrono::duration_cast<std::chrono::milliseconds>(b - a).count(); };

    kg_context* ctx = nullptr;
    kg_table* table = nullptr;
    kg_functions* fn = nullptr;
    kg_fasta* fa = nullptr;
    kg_result* res = nullptr;
    int rc = 1;
    do {
        if (kg_init(device, &ctx) != KG_OK) break;
        if (kg_functions_load(dir, &fn) != KG_OK) break; // KGJ:759
        auto t0 = now();
        // The cache is trusted only if it was built from the kmer.table.mem_map that is in -D now (size + mtime in its
        // header) and its body checksum still holds; otherwise the table is rebuilt from the reference-format file.  A
        // data directory that holds no table file at all (a deployment that ships only the cache) uses the cache as it is.
        bool have = false;
        if (cache) {
            struct stat sst;
            const std::string base = std::string(dir) + "/kmer.table.mem_map";
            const bool source = stat((base + ".gz").c_str(), &sst) == 0 || stat(base.c_str(), &sst) == 0;
            have = (source ? kg_table_load_cached_checked(ctx, cache, dir, &table) : kg_table_load_cached(ctx, cache, &table)) == KG_OK;
            if (!have && access(cache, F_OK) == 0) fprintf(stderr, "Note: table cache not used: %s\n", kg_last_error());
        }
        if (!have) {
            if (kg_table_load(ctx, dir, &table) != KG_OK) break; // KGJ:774
            if (cache && kg_table_save(ctx, table, cache) != KG_OK) fprintf(stderr, "Warning: %s\n", kg_last_error());
        }
        info("Table load time: " + std::to_string(ms(t0, now())) + " ms.");
        const int rflags = (debug ? 1 : 0) | (dna_ranges ? 2 : 0);
        if (batch_mb > 0) { // streaming: batch i+1 is read and parsed while batch i is on the GPU and in the report writer
            kg_fasta_stream* fs = nullptr;
            kg_report* rep = nullptr;
            if (kg_fasta_stream_open(query, (size_t)batch_mb << 20, &fs) != KG_OK) break;
            if (kg_report_open(outp, &rep) != KG_OK) {
                kg_fasta_stream_close(fs);
                break;
            }
            auto t1 = now();
            kg_fasta* nextb = nullptr;
            int nrc = kg_fasta_stream_next(fs, &nextb);
            std::string nerr = nrc != KG_OK ? kg_last_error() : "";
            bool ok = nrc == KG_OK;
            size_t nbatches = 0;
            while (ok && nextb) {
                fa = nextb;
                nextb = nullptr;
                std::thread reader([&] {
                    nrc = kg_fasta_stream_next(fs, &nextb);
                    if (nrc != KG_OK) nerr = kg_last_error(); // thread-local message: carried over by hand
                });
                ok = kg_run(ctx, table, aa ? KG_MODE_AA : KG_MODE_DNA, kg_fasta_bytes(fa), kg_fasta_offsets(fa), kg_fasta_count(fa), &prm, &res) == KG_OK &&
                     kg_report_add(rep, aa ? KG_MODE_AA : KG_MODE_DNA, rflags, fa, fn, table, res) == KG_OK;
                reader.join();
                kg_result_free(res);
                res = nullptr;
                kg_fasta_free(fa);
                fa = nullptr;
                nbatches++;
                if (ok && nrc != KG_OK) {
                    kg_set_error("%s", nerr.c_str());
                    ok = false;
                }
            }
            if (!ok && nrc != KG_OK && nbatches == 0) kg_set_error("%s", nerr.c_str());
            kg_fasta_free(nextb);
            kg_fasta_stream_close(fs);
            const bool closed = kg_report_close(rep) == KG_OK;
            if (!ok || !closed) break;
            info("Streaming time (" + std::to_string(nbatches) + " batches): " + std::to_string(ms(t1, now())) + " ms.");
            rc = 0;
            break;
        }
        auto t1 = now();
        if (kg_fasta_read(query, &fa) != KG_OK) break; // KGJ:778
        info("Preparation time: " + std::to_string(ms(t1, now())) + " ms.");
        auto t2 = now();
        if (kg_run(ctx, table, aa ? KG_MODE_AA : KG_MODE_DNA, kg_fasta_bytes(fa), kg_fasta_offsets(fa), kg_fasta_count(fa), &prm, &res) != KG_OK) break;
        info("Lookup time: " + std::to_string(ms(t2, now())) + " ms.");
        auto t3 = now();
        if (kg_report_write(outp, aa ? KG_MODE_AA : KG_MODE_DNA, rflags, fa, fn, table, res) != KG_OK) break;
        info("Grouping time: " + std::to_string(ms(t3, now())) + " ms.");
        rc = 0;
    } while (0);
    if (rc) fprintf(stderr, "Error: %s\n", kg_last_error());
    kg_result_free(res);
    kg_fasta_free(fa);
    kg_functions_free(fn);
    kg_table_free(table);
    kg_shutdown(ctx);
    return rc;
}
