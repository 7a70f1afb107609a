// writers.hpp — output files of `genmap map`: raw (.map/.freq8/.freq16), .txt, .wig + .chrom.sizes,
// .bedgraph, .bed.  Byte-identical to the reference's writers (src/output.hpp:10-187, dispatch and verbose
// lines src/mappability.hpp:69-155); structured as one run-length pass feeding small format emitters.
#pragma once
#include <fcntl.h>
#include <sys/time.h>
#include <unistd.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <iostream>
#include <string>
#include <thread>
#include <vector>

namespace gmbcli {

enum class OutputType { mappability, frequency_small, frequency_large };

struct Outputs { bool raw, txt, wig, bedgraph, bed, verbose; unsigned threads = 1; };

inline double now_s()
{
    struct timeval t;
    gettimeofday(&t, nullptr);
    return t.tv_sec + t.tv_usec * 1e-6;
}

// buffered FILE* writer; numbers formatted like std::ostream's defaults (%g for float: 6 significant digits)
class Sink {
public:
    explicit Sink(const std::string& path) : f_(std::fopen(path.c_str(), "wb")) { if (f_) std::setvbuf(f_, nullptr, _IOFBF, 1 << 20); }
    ~Sink() { if (f_) std::fclose(f_); }
    bool ok() const { return f_ != nullptr; }
    void str(const std::string& s) { std::fwrite(s.data(), 1, s.size(), f_); }
    void ch(char c) { std::fputc(c, f_); }
    void u64(uint64_t v)
    {
        char buf[24];
        int n = 0;
        do { buf[n++] = (char)('0' + v % 10); v /= 10; } while (v);
        while (n) std::fputc(buf[--n], f_);
    }
    void flt(float v) { std::fprintf(f_, "%g", (double)v); }
    // 1/v as the reference prints it (operator<< of a float: %g), from a table: only 65536 different values exist
    void inverse(uint32_t v)
    {
        if (v >= (1u << 16)) { flt(1.0f / static_cast<float>(v)); return; }
        const char* t = inverse_text(v);
        std::fwrite(t + 1, 1, (size_t)t[0], f_);
    }
    static const char* inverse_text(uint32_t v) // 16 bytes per value: length, then the characters
    {
        struct Table {
            std::vector<char> text;
            Table() : text(16u << 16)
            {
                for (uint32_t x = 0; x < (1u << 16); ++x) {
                    const float f = x != 0 ? 1.0f / static_cast<float>(x) : 0.0f;
                    text[16 * x] = (char)std::snprintf(&text[16 * x + 1], 15, "%g", (double)f);
                }
            }
        };
        static const Table table;
        return &table.text[16 * v];
    }
    void bytes(const void* p, size_t n) { std::fwrite(p, 1, n, f_); }
    template <class T>
    void value(T v, bool mappability) // a frequency, or its inverse as float (0 stays 0)
    {
        if (mappability) inverse((uint32_t)v);
        else u64(v);
    }
private:
    FILE* f_;
};

struct Run { uint64_t start, len; uint32_t value; };

// maximal runs of equal values inside one sequence [begin, end)
template <class T, class F>
void for_each_run(const T* c, uint64_t begin, uint64_t end, F&& f)
{
    uint64_t i = begin;
    while (i < end) {
        uint64_t j = i + 1;
        while (j < end && c[j] == c[i]) ++j;
        f(Run{i - begin, j - i, (uint32_t)c[i]});
        i = j;
    }
}

// Where the track writers get their runs from: the frequency vector in host memory (scanned here), or the run
// list the GPU produced (gmb_map_runs: only the runs were copied to the host).
template <class T>
struct VectorRuns {
    const T* c;
    const std::vector<uint64_t>& cum; // cumulative sequence lengths, first = 0
    template <class F> void runs(size_t s, F&& f) const { for_each_run(c, cum[s], cum[s + 1], f); }
};

struct ListRuns {
    const uint64_t* start; // ascending file-local run starts; a new run at every sequence start
    const uint16_t* value;
    uint64_t n;
    const std::vector<uint64_t>& cum;
    template <class F> void runs(size_t s, F&& f) const
    {
        uint64_t r = std::lower_bound(start, start + n, cum[s]) - start;
        for (; r < n && start[r] < cum[s + 1]; ++r) f(at(r, s));
    }
    // run indices [first, last) of sequence s, and run r of that sequence
    std::pair<uint64_t, uint64_t> range(size_t s) const
    {
        return {(uint64_t)(std::lower_bound(start, start + n, cum[s]) - start), (uint64_t)(std::lower_bound(start, start + n, cum[s + 1]) - start)};
    }
    Run at(uint64_t r, size_t s) const
    {
        const uint64_t end = r + 1 < n ? std::min(start[r + 1], cum[s + 1]) : cum[s + 1];
        return Run{start[r] - cum[s], end - start[r], value[r]};
    }
};

// text building blocks of the parallel run formatters
inline void append_u64(std::string& out, uint64_t v)
{
    char buf[24];
    int n = 0;
    do { buf[n++] = (char)('0' + v % 10); v /= 10; } while (v);
    while (n) out.push_back(buf[--n]);
}
inline void append_value(std::string& out, uint32_t v, bool mappability)
{
    if (!mappability) { append_u64(out, v); return; }
    if (v >= (1u << 16)) { char b[32]; out.append(b, (size_t)std::snprintf(b, sizeof b, "%g", (double)(1.0f / static_cast<float>(v)))); return; }
    const char* t = Sink::inverse_text(v);
    out.append(t + 1, (size_t)t[0]);
}

// Formats the runs of every sequence with `threads` workers (consecutive chunks of runs, written in order).
// emit(run, last_span, seq, out): appends the text of one run; last_span = length of the previous run of the
// sequence that was written (value != 0), 0 if none — the only state the wig format carries from run to run.
template <class Emit>
void format_runs_parallel(const ListRuns& src, size_t n_seq, unsigned threads, Sink& o, Emit&& emit)
{
    if (threads < 1) threads = 1;
    const uint64_t chunk = 1u << 18; // runs per work item
    std::vector<std::string> bufs(threads);
    for (size_t s = 0; s < n_seq; ++s) {
        const std::pair<uint64_t, uint64_t> rr = src.range(s);
        for (uint64_t at = rr.first; at < rr.second; at += chunk * threads) {
            std::vector<std::thread> workers;
            unsigned used = 0;
            for (unsigned t = 0; t < threads && at + t * chunk < rr.second; ++t, ++used) {
                const uint64_t b = at + t * chunk, e = std::min(rr.second, b + chunk);
                auto work = [&, t, b, e, s] {
                    uint64_t last_span = 0;
                    for (uint64_t r = b; r-- > rr.first;) // the nearest earlier run of this sequence that was written
                        if (src.value[r] != 0) { last_span = src.at(r, s).len; break; }
                    std::string& out = bufs[t];
                    out.clear();
                    for (uint64_t r = b; r < e; ++r) {
                        const Run run = src.at(r, s);
                        if (run.value == 0) continue;
                        emit(run, last_span, s, out);
                        last_span = run.len;
                    }
                };
                if (threads == 1) work(); else workers.emplace_back(work);
            }
            for (std::thread& w : workers) w.join();
            for (unsigned t = 0; t < used; ++t) o.str(bufs[t]);
        }
    }
}

inline std::vector<uint64_t> cumulative(const std::vector<uint64_t>& lens)
{
    std::vector<uint64_t> cum{0};
    for (uint64_t l : lens) cum.push_back(cum.back() + l);
    return cum;
}

// src/output.hpp:10-31: the vector as it is (.freq8 / .freq16) or its inverses as floats (.map).  One file, written
// by `threads` workers at their own offsets (pwrite): at 3 Gbp the raw file has 6-12 GB and a single buffered writer was
// the longest step of a whole `genmap map` run.
template <class T>
void write_raw(const T* c, uint64_t n, const std::string& path, bool mappability, unsigned threads = 1)
{
    const int fd = ::open(path.c_str(), O_WRONLY | O_CREAT | O_TRUNC, 0644);
    if (fd < 0) { std::cerr << "ERROR: cannot write " << path << "\n"; return; }
    const size_t elem = mappability ? sizeof(float) : sizeof(T);
    const uint64_t piece = 8u << 20; // values per piece
    const uint64_t n_pieces = (n + piece - 1) / piece;
    if (threads < 1) threads = 1;
    if (threads > n_pieces) threads = (unsigned)std::max<uint64_t>(1, n_pieces);
    std::vector<char> failed(threads, 0);
    auto work = [&](unsigned t) {
        std::vector<float> buf(mappability ? piece : 0);
        for (uint64_t q = t; q < n_pieces; q += threads) {
            const uint64_t b = q * piece, e = std::min(n, b + piece);
            const char* src = reinterpret_cast<const char*>(c + b);
            if (mappability) { // src/output.hpp:17-24
                for (uint64_t i = b; i < e; ++i) buf[i - b] = c[i] != 0 ? 1.0f / static_cast<float>(c[i]) : 0.0f;
                src = reinterpret_cast<const char*>(buf.data());
            }
            size_t left = (e - b) * elem;
            off_t at = (off_t)(b * elem);
            while (left) {
                const ssize_t w = ::pwrite(fd, src, left, at);
                if (w <= 0) { failed[t] = 1; return; }
                src += w; at += w; left -= (size_t)w;
            }
        }
    };
    if (threads == 1) work(0);
    else {
        std::vector<std::thread> pool;
        for (unsigned t = 0; t < threads; ++t) pool.emplace_back(work, t);
        for (std::thread& w : pool) w.join();
    }
    if (::close(fd) != 0 || std::count(failed.begin(), failed.end(), 1)) std::cerr << "ERROR: short write to " << path << "\n";
}

// The runs of a vector in host memory as a run list (what gmb_map_runs returns from the device), found by `threads`
// workers on consecutive chunks; a run that continues across a chunk border is one run, and a new run starts at every
// sequence start.  Lets the threaded run formatters serve the track formats also when the vector itself was asked for.
template <class T>
void runs_of_vector(const T* c, const std::vector<uint64_t>& cum, unsigned threads, std::vector<uint64_t>& start, std::vector<uint16_t>& value)
{
    const uint64_t n = cum.back();
    if (threads < 1) threads = 1;
    const uint64_t chunk = std::max<uint64_t>(1u << 20, (n + threads - 1) / threads);
    const uint64_t n_chunks = n ? (n + chunk - 1) / chunk : 0;
    std::vector<std::vector<uint64_t>> st(n_chunks);
    std::vector<std::vector<uint16_t>> va(n_chunks);
    auto work = [&](uint64_t q) {
        const uint64_t b = q * chunk, e = std::min(n, b + chunk);
        size_t s = std::upper_bound(cum.begin(), cum.end(), b) - cum.begin(); // next sequence start after b
        for (uint64_t i = b; i < e; ++i) {
            bool head = i == b || c[i] != c[i - 1];
            while (s < cum.size() && cum[s] < i) ++s;
            if (s < cum.size() && cum[s] == i) { head = true; ++s; }
            if (head) { st[q].push_back(i); va[q].push_back((uint16_t)c[i]); }
        }
    };
    std::vector<std::thread> pool;
    for (uint64_t q = 0; q < n_chunks; ++q) {
        if (threads == 1) work(q);
        else pool.emplace_back(work, q);
    }
    for (std::thread& w : pool) w.join();
    start.clear(); value.clear();
    for (uint64_t q = 0; q < n_chunks; ++q)
        for (size_t r = 0; r < st[q].size(); ++r) {
            const uint64_t at = st[q][r];
            if (r == 0 && q != 0 && !value.empty() && value.back() == va[q][r] && !std::binary_search(cum.begin(), cum.end(), at)) continue;
            start.push_back(at); value.push_back(va[q][r]);
        }
}

// values [begin, end) of one sequence as text, separated by single spaces (no leading / trailing separator)
template <class T>
void format_values(const T* c, uint64_t begin, uint64_t end, bool mappability, std::string& out)
{
    out.clear();
    out.reserve((end - begin) * (mappability ? 6 : 3));
    char buf[24];
    for (uint64_t i = begin; i < end; ++i) {
        if (i != begin) out.push_back(' ');
        const uint32_t v = (uint32_t)c[i];
        if (mappability) {
            const char* t = Sink::inverse_text(v);
            out.append(t + 1, (size_t)t[0]);
        } else {
            int n = 0;
            uint32_t x = v;
            do { buf[n++] = (char)('0' + x % 10); x /= 10; } while (x);
            while (n) out.push_back(buf[--n]);
        }
    }
}

// src/output.hpp:40-69: '>' name, then the values of the sequence separated by single spaces.  Formatting is the
// cost (one number per position): `threads` workers format consecutive chunks, the file is written in order.
template <class T>
void write_txt(const T* c, const std::string& prefix, const std::vector<std::string>& names,
               const std::vector<uint64_t>& lens, bool mappability, unsigned threads = 1)
{
    Sink o(prefix + ".txt");
    if (!o.ok()) { std::cerr << "ERROR: cannot write " << prefix << ".txt\n"; return; }
    if (threads < 1) threads = 1;
    const uint64_t chunk = 2u << 20;
    std::vector<std::string> bufs(threads);
    uint64_t begin = 0;
    for (size_t s = 0; s < lens.size(); ++s) {
        o.ch('>'); o.str(names[s]); o.ch('\n');
        for (uint64_t at = 0; at < lens[s]; at += chunk * threads) {
            std::vector<std::thread> workers;
            unsigned used = 0;
            for (unsigned t = 0; t < threads && at + t * chunk < lens[s]; ++t, ++used) {
                const uint64_t b = begin + at + t * chunk, e = std::min(begin + lens[s], b + chunk);
                if (threads == 1) format_values(c, b, e, mappability, bufs[t]);
                else workers.emplace_back([&, t, b, e] { format_values(c, b, e, mappability, bufs[t]); });
            }
            for (std::thread& w : workers) w.join();
            for (unsigned t = 0; t < used; ++t) {
                if (at + t * chunk) o.ch(' ');
                o.str(bufs[t]);
            }
        }
        o.ch('\n');
        begin += lens[s];
    }
}

template <class Source>
void write_wig(const Source& src, const std::string& prefix, const std::vector<std::string>& names,
               const std::vector<uint64_t>& lens, bool mappability)
{
    {
        Sink o(prefix + ".wig");
        if (!o.ok()) { std::cerr << "ERROR: cannot write " << prefix << ".wig\n"; return; }
        for (size_t s = 0; s < lens.size(); ++s) {
            uint64_t last_span = 0; // a new variableStep header whenever the run length changes (:96-99); reset per sequence
            src.runs(s, [&](const Run& r) {
                if (r.value == 0) return; // zero runs are skipped (:96)
                if (last_span != r.len) {
                    o.str("variableStep chrom="); o.str(names[s]); o.str(" span="); o.u64(r.len); o.ch('\n');
                }
                o.u64(r.start + 1); o.ch(' '); o.value(r.value, mappability); o.ch('\n'); // positions start at 1
                last_span = r.len;
            });
        }
    }
    Sink cs(prefix + ".chrom.sizes");
    if (!cs.ok()) return;
    for (size_t s = 0; s < lens.size(); ++s) { cs.str(names[s]); cs.ch('\t'); cs.u64(lens[s]); cs.ch('\n'); }
}

template <class Source>
void write_bedgraph(const Source& src, const std::string& prefix, const std::vector<std::string>& names,
                    const std::vector<uint64_t>& lens, bool bedgraph_format, bool mappability)
{
    Sink o(prefix + (bedgraph_format ? ".bedgraph" : ".bed"));
    if (!o.ok()) { std::cerr << "ERROR: cannot write " << prefix << (bedgraph_format ? ".bedgraph" : ".bed") << "\n"; return; }
    for (size_t s = 0; s < lens.size(); ++s) {
        src.runs(s, [&](const Run& r) {
            if (r.value == 0) return; // src/output.hpp:157
            o.str(names[s]); o.ch('\t'); o.u64(r.start); o.ch('\t'); o.u64(r.start + r.len); o.ch('\t');
            if (!bedgraph_format) { o.ch('-'); o.ch('\t'); }
            o.value(r.value, mappability); o.ch('\n');
        });
    }
}

inline void write_track_outputs(const ListRuns& src, const std::string& prefix, const std::vector<std::string>& names,
                                const std::vector<uint64_t>& lens, OutputType type, const Outputs& o);

template <class T>
void write_outputs(const T* c, uint64_t n, const std::string& prefix, const std::vector<std::string>& names,
                   const std::vector<uint64_t>& lens, OutputType type, const Outputs& o)
{
    const bool mapp = type == OutputType::mappability;
    auto timed = [&](const char* what, auto&& fn) {
        const double t0 = now_s();
        fn();
        if (o.verbose) std::cout << "- " << what << " written in " << (std::round((now_s() - t0) * 100.0) / 100.0) << " seconds\n";
    };
    if (o.raw) timed("RAW file", [&] {
        write_raw(c, n, prefix + (mapp ? ".map" : type == OutputType::frequency_small ? ".freq8" : ".freq16"), mapp, o.threads);
    });
    if (o.txt) timed("TXT file", [&] { write_txt(c, prefix, names, lens, mapp, o.threads); });
    if (!(o.wig || o.bedgraph || o.bed)) return;
    const std::vector<uint64_t> cum = cumulative(lens);
    if (o.threads <= 1) { // one scan per format, as the reference does it
        const VectorRuns<T> src{c, cum};
        if (o.wig) timed("WIG file", [&] { write_wig(src, prefix, names, lens, mapp); });
        if (o.bedgraph) timed("bedgraph file", [&] { write_bedgraph(src, prefix, names, lens, true, mapp); });
        if (o.bed) timed("BED file", [&] { write_bedgraph(src, prefix, names, lens, false, mapp); });
        return;
    }
    // several host threads: the runs once (threaded scan), then the threaded run formatters
    std::vector<uint64_t> run_start;
    std::vector<uint16_t> run_value;
    runs_of_vector(c, cum, o.threads, run_start, run_value);
    write_track_outputs(ListRuns{run_start.data(), run_value.data(), run_start.size(), cum}, prefix, names, lens, type, o);
}

// run-list overloads: the same files, formatted by several host threads
inline void write_wig_runs(const ListRuns& src, const std::string& prefix, const std::vector<std::string>& names,
                           const std::vector<uint64_t>& lens, bool mappability, unsigned threads)
{
    {
        Sink o(prefix + ".wig");
        if (!o.ok()) { std::cerr << "ERROR: cannot write " << prefix << ".wig\n"; return; }
        format_runs_parallel(src, lens.size(), threads, o, [&](const Run& r, uint64_t last_span, size_t s, std::string& out) {
            if (last_span != r.len) { // src/output.hpp:96-99
                out += "variableStep chrom="; out += names[s]; out += " span="; append_u64(out, r.len); out.push_back('\n');
            }
            append_u64(out, r.start + 1); out.push_back(' '); append_value(out, r.value, mappability); out.push_back('\n');
        });
    }
    Sink cs(prefix + ".chrom.sizes");
    if (!cs.ok()) return;
    for (size_t s = 0; s < lens.size(); ++s) { cs.str(names[s]); cs.ch('\t'); cs.u64(lens[s]); cs.ch('\n'); }
}

inline void write_bedgraph_runs(const ListRuns& src, const std::string& prefix, const std::vector<std::string>& names,
                                const std::vector<uint64_t>& lens, bool bedgraph_format, bool mappability, unsigned threads)
{
    Sink o(prefix + (bedgraph_format ? ".bedgraph" : ".bed"));
    if (!o.ok()) { std::cerr << "ERROR: cannot write " << prefix << (bedgraph_format ? ".bedgraph" : ".bed") << "\n"; return; }
    format_runs_parallel(src, lens.size(), threads, o, [&](const Run& r, uint64_t, size_t s, std::string& out) {
        out += names[s]; out.push_back('\t'); append_u64(out, r.start); out.push_back('\t'); append_u64(out, r.start + r.len); out.push_back('\t');
        if (!bedgraph_format) { out.push_back('-'); out.push_back('\t'); }
        append_value(out, r.value, mappability); out.push_back('\n');
    });
}

// the track formats from a run list (no frequency vector on the host)
inline void write_track_outputs(const ListRuns& src, const std::string& prefix, const std::vector<std::string>& names,
                                const std::vector<uint64_t>& lens, OutputType type, const Outputs& o)
{
    const bool mapp = type == OutputType::mappability;
    auto timed = [&](const char* what, auto&& fn) {
        const double t0 = now_s();
        fn();
        if (o.verbose) std::cout << "- " << what << " written in " << (std::round((now_s() - t0) * 100.0) / 100.0) << " seconds\n";
    };
    if (o.wig) timed("WIG file", [&] { write_wig_runs(src, prefix, names, lens, mapp, o.threads); });
    if (o.bedgraph) timed("bedgraph file", [&] { write_bedgraph_runs(src, prefix, names, lens, true, mapp, o.threads); });
    if (o.bed) timed("BED file", [&] { write_bedgraph_runs(src, prefix, names, lens, false, mapp, o.threads); });
}

} // namespace gmbcli
