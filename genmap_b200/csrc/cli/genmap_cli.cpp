// genmap_cli.cpp — the `genmap` command line (index / map) on top of the C ABI.
//
// Keeps the reference's user surface for the map step — flag spellings, messages, exit codes, output
// naming (src/genmap.cpp:16-94, src/indexing.hpp:277-510, src/mappability.hpp:409-642) — while the
// per-position work happens in libgenmap_b200.so.  Thin host plumbing: no arithmetic lives here.
#include <sys/stat.h>
#include <sys/time.h>
#include <dirent.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <map>
#include <memory>
#include <set>
#include <sstream>
#include <string>
#include <thread>
#include <vector>

#include "../../../include/genmap_b200.h"
#include "writers.hpp"

namespace {

const char* kVersion = "1.3.0-b200";

double wall()
{
    struct timeval t;
    gettimeofday(&t, nullptr);
    return t.tv_sec + t.tv_usec * 1e-6;
}

double round2(double x) { return std::round(x * 100.0) / 100.0; }

bool is_dir(const std::string& p)
{
    struct stat st;
    return stat(p.c_str(), &st) == 0 && S_ISDIR(st.st_mode);
}
bool exists(const std::string& p)
{
    struct stat st;
    return stat(p.c_str(), &st) == 0;
}

// ---- a small option parser with SeqAn ArgumentParser's spellings (-K / --length, multi-letter -nc) ----
struct OptSpec { std::string s, l; bool has_value; };
struct Args {
    std::map<std::string, std::string> val; // keyed by long name
    std::set<std::string> flag;
    bool has(const std::string& k) const { return val.count(k) || flag.count(k); }
};

// returns 0 ok, 1 error (message printed), 2 help/version printed
int parse_args(const std::string& prog, const std::vector<OptSpec>& specs, int argc, char const** argv, Args& out,
               const std::string& help_text)
{
    for (int i = 1; i < argc; ++i) {
        std::string a = argv[i];
        if (a == "-h" || a == "--help") { std::cout << help_text; return 2; }
        if (a == "--version") { std::cout << prog << " version: " << kVersion << "\n"; return 2; }
        if (a == "--copyright") { std::cout << "genmap-b200: B200-native (k,e)-mappability; CLI surface after GenMap (3-clause BSD).\n"; return 2; }
        const OptSpec* sp = nullptr;
        std::string inline_val;
        bool has_inline = false;
        for (const OptSpec& s : specs) {
            if (a == "-" + s.s || a == "--" + s.l) { sp = &s; break; }
            if (a.rfind("--" + s.l + "=", 0) == 0) { sp = &s; inline_val = a.substr(s.l.size() + 3); has_inline = true; break; }
        }
        if (!sp) {
            std::cerr << prog << ": Unknown option " << a << "\n";
            return 1;
        }
        if (!sp->has_value) { out.flag.insert(sp->l); continue; }
        if (!has_inline) {
            if (i + 1 >= argc) { std::cerr << prog << ": Missing value for option: -" << sp->s << ", --" << sp->l << "\n"; return 1; }
            inline_val = argv[++i];
        }
        out.val[sp->l] = inline_val;
    }
    return 0;
}

bool to_uint(const std::string& s, uint64_t& v)
{
    if (s.empty()) return false;
    char* end = nullptr;
    errno = 0;
    unsigned long long x = std::strtoull(s.c_str(), &end, 10);
    if (errno || *end || s[0] == '-') return false;
    v = x;
    return true;
}

// ---- FASTA ---------------------------------------------------------------------------------------------
struct Record { std::string id; uint64_t length; };

int8_t code_of(unsigned char c)
{
    switch (c) {
        case 'A': case 'a': return 0;
        case 'C': case 'c': return 1;
        case 'G': case 'g': return 2;
        case 'T': case 't': case 'U': case 'u': return 3;
        default: return 4; // everything else becomes N (src/indexing.hpp:13-20)
    }
}

// reads one FASTA file: appends codes/limits, records (ids cut at the first whitespace if still unique,
// empty records skipped: src/indexing.hpp:208-275)
bool read_fasta(const std::string& path, std::vector<uint8_t>& codes, std::vector<uint64_t>& limits,
                std::vector<Record>& recs)
{
    std::ifstream in(path, std::ios::binary);
    if (!in) return false;
    std::vector<std::string> ids;
    std::vector<uint64_t> lens;
    std::string line, cur_id;
    bool have = false;
    uint64_t cur_len = 0;
    auto flush = [&]() {
        if (have && cur_len > 0) { ids.push_back(cur_id); lens.push_back(cur_len); limits.push_back(codes.size()); }
        else if (have) { /* empty record: nothing was appended */ }
        cur_len = 0;
    };
    // FASTQ (the reference's SeqFileIn detects the format from the first character; -FD accepts *.fastq): '@id',
    // sequence lines up to the '+' line, then as many quality characters as the record has bases (ignored)
    const int first = in.peek();
    if (first == '@') {
        while (std::getline(in, line)) {
            if (!line.empty() && line.back() == '\r') line.pop_back();
            if (line.empty()) continue;
            if (line[0] != '@') return false; // malformed record
            have = true;
            cur_id = line.substr(1);
            while (std::getline(in, line)) {
                if (!line.empty() && line.back() == '\r') line.pop_back();
                if (!line.empty() && line[0] == '+') break;
                for (unsigned char ch : line) {
                    if (ch == ' ' || ch == '\t') continue;
                    codes.push_back((uint8_t)code_of(ch));
                    ++cur_len;
                }
            }
            uint64_t quals = 0;
            while (quals < cur_len && std::getline(in, line)) {
                if (!line.empty() && line.back() == '\r') line.pop_back();
                quals += line.size();
            }
            flush();
        }
        have = false;
    }
    if (first != '@') {
        // FASTA: the file in 16 MB pieces through a 256-entry code table (3 GB of sequence: a getline per 80-column line and
        // a push_back per base were 15 of the 21 s of `genmap index` at 3 Gbp)
        static uint8_t lut[256];
        static bool lut_ready = false;
        if (!lut_ready) {
            for (int c = 0; c < 256; ++c) lut[c] = (uint8_t)code_of((unsigned char)c);
            lut[(unsigned char)' '] = lut[(unsigned char)'\t'] = lut[(unsigned char)'\r'] = 0xff; // skipped inside sequence lines
            lut_ready = true;
        }
        in.seekg(0, std::ios::end);
        const std::streamoff file_size = in.tellg();
        in.seekg(0, std::ios::beg);
        if (file_size > 0) codes.reserve(codes.size() + (size_t)file_size);
        std::vector<char> buf(16u << 20);
        bool in_header = false, at_line_start = true;
        while (in.read(buf.data(), (std::streamsize)buf.size()) || in.gcount() > 0) {
            const size_t n = (size_t)in.gcount();
            size_t i = 0;
            while (i < n) {
                if (in_header) { // the rest of the '>' line is the record's id
                    const char* nl = static_cast<const char*>(std::memchr(buf.data() + i, '\n', n - i));
                    const size_t e = nl ? (size_t)(nl - buf.data()) : n;
                    cur_id.append(buf.data() + i, e - i);
                    i = e;
                    if (nl) { in_header = false; at_line_start = true; ++i; if (!cur_id.empty() && cur_id.back() == '\r') cur_id.pop_back(); }
                    continue;
                }
                const unsigned char ch = (unsigned char)buf[i];
                if (ch == '\n') { at_line_start = true; ++i; continue; }
                if (at_line_start && ch == '>') {
                    flush();
                    have = true;
                    cur_id.clear();
                    in_header = true;
                    at_line_start = false;
                    ++i;
                    continue;
                }
                at_line_start = false;
                if (have) { // a run of sequence characters up to the end of the line or of the piece
                    const size_t start = codes.size();
                    size_t j = i;
                    while (j < n && buf[j] != '\n') ++j;
                    codes.resize(start + (j - i));
                    uint8_t* out = codes.data() + start;
                    size_t m = 0;
                    for (size_t k = i; k < j; ++k) {
                        const uint8_t c = lut[(unsigned char)buf[k]];
                        out[m] = c;
                        m += c != 0xff;
                    }
                    codes.resize(start + m);
                    cur_len += m;
                    i = j;
                } else {
                    ++i; // text before the first record
                }
            }
        }
    }
    flush();
    std::vector<std::string> shortened;
    for (const std::string& id : ids) {
        size_t p = 0;
        while (p < id.size() && !std::isspace((unsigned char)id[p])) ++p;
        shortened.push_back(id.substr(0, p));
    }
    std::vector<std::string> sorted = shortened;
    std::sort(sorted.begin(), sorted.end());
    const bool unique = std::adjacent_find(sorted.begin(), sorted.end()) == sorted.end();
    for (size_t i = 0; i < ids.size(); ++i) recs.push_back(Record{unique ? shortened[i] : ids[i], lens[i]});
    return true;
}

std::string file_name_of(const std::string& path)
{
    size_t p = path.find_last_of('/');
    return p == std::string::npos ? path : path.substr(p + 1);
}

bool has_fasta_ext(const std::string& name)
{
    static const char* exts[] = {"fsa", "fna", "fastq", "fasta", "fas", "faa", "fa"};
    size_t p = name.find_last_of('.');
    if (p == std::string::npos) return false;
    std::string e = name.substr(p + 1);
    for (const char* x : exts) if (e == x) return true;
    return false;
}

// ---- index ---------------------------------------------------------------------------------------------
const char* kIndexHelp =
    "GenMap index (B200 build)\n\n"
    "    genmap index -F genome.fa | -FD fasta_dir  -I index_dir [-v]\n\n"
    "  -F,  --fasta-file       Path to the fasta file.\n"
    "  -FD, --fasta-directory  Path to the directory of fasta files (.fsa .fna .fastq .fasta .fas .faa .fa).\n"
    "  -I,  --index            Path to the index (directory must not exist yet).\n"
    "  -A,  --algorithm        accepted for compatibility (divsufsort|skew); the suffix array is built on the\n"
    "                          GPU (prefix doubling) or, with --host-build / without a GPU, by host SA-IS.\n"
    "  -S,  --sampling         accepted for compatibility: the index stores the FULL suffix array (4 bytes per\n"
    "                          base; used by --exclude-pseudo), unless --no-sa is given.\n"
    "  -xn, --no-sa            do not store the suffix array (smaller index; --exclude-pseudo unavailable).\n"
    "  -xf, --reference-format also write the index in the reference's own on-disk format (SeqAn fibres, suffix array\n"
    "                          sampled with -S, default 10): the directory then opens in the original genmap as well.\n"
    "  -v,  --verbose\n";

int index_main(int argc, char const** argv)
{
    std::vector<OptSpec> specs = {{"F", "fasta-file", true}, {"FD", "fasta-directory", true}, {"I", "index", true},
                                  {"A", "algorithm", true}, {"S", "sampling", true}, {"v", "verbose", false},
                                  {"xa", "seqno", true}, {"xb", "seqpos", true}, {"xc", "bwtlen", true},
                                  {"xn", "no-sa", false}, {"xh", "host-build", false}, {"xf", "reference-format", false}};
    Args a;
    int rc = parse_args("GenMap index", specs, argc, argv, a, kIndexHelp);
    if (rc == 2) return 0;
    if (rc) return 1;
    if (!a.has("index")) { std::cerr << "GenMap index: Missing value for option: -I, --index\n"; return 1; }
    const bool f = a.has("fasta-file"), fd = a.has("fasta-directory");
    if (f && fd) { std::cerr << "ERROR: You can only use eiher --fasta-file or --fasta-directory, not both.\n"; return 1; }
    if (!f && !fd) { std::cerr << "ERROR: You forgot to specify --fasta-file or --fasta-directory.\n"; return 1; }
    if (a.has("algorithm")) {
        std::string al = a.val["algorithm"];
        std::transform(al.begin(), al.end(), al.begin(), ::tolower);
        if (al != "divsufsort" && al != "skew") { std::cerr << "GenMap index: the given value '" << a.val["algorithm"] << "' is not in the list of allowed values [divsufsort, skew]\n"; return 1; }
    }
    const std::string index_dir = a.val["index"];
    std::vector<std::pair<std::string, std::string>> files; // {full path, file name}
    if (fd) {
        std::string dir = a.val["fasta-directory"];
        if (!is_dir(dir)) { std::cerr << "ERROR: The fasta directory does not exist!\n"; return 1; }
        if (dir.back() != '/') dir += '/';
        DIR* d = opendir(dir.c_str());
        if (d) {
            while (dirent* e = readdir(d)) {
                std::string n = e->d_name;
                if (has_fasta_ext(n) && !is_dir(dir + n)) files.push_back({dir + n, n});
            }
            closedir(d);
        }
        std::sort(files.begin(), files.end(), [](auto const& x, auto const& y) { return x.second < y.second; }); // src/indexing.hpp:407
    } else {
        const std::string p = a.val["fasta-file"];
        if (!exists(p) || is_dir(p)) { std::cerr << "ERROR: The fasta file does not exist!\n"; return 1; }
        files.push_back({p, file_name_of(p)});
    }
    if (exists(index_dir)) {
        std::cerr << "ERROR: The directory for the index already exists at " << index_dir << "\n"
                  << "       Please remove it, or choose a different location.\n";
        return 1;
    }
    if (mkdir(index_dir.c_str(), 0755)) { std::cerr << "ERROR: Cannot create directory at " << index_dir << "\n"; return 1; }

    std::vector<uint8_t> codes;
    std::vector<uint64_t> limits{0};
    std::vector<std::string> ids_lines;
    for (auto const& file : files) {
        std::vector<Record> recs;
        if (!read_fasta(file.first, codes, limits, recs)) { rmdir(index_dir.c_str()); std::cerr << "ERROR: cannot read " << file.first << "\n"; return 1; }
        if (recs.empty()) std::cerr << "WARNING: The fasta file " << file.first << " seems to be empty. Excluded from indexing.\n";
        for (const Record& r : recs) ids_lines.push_back(file.second + ";" + std::to_string(r.length) + ";" + r.id);
    }
    if (fd) {
        if (ids_lines.empty()) { rmdir(index_dir.c_str()); std::cerr << "ERROR: No (non-empty) fasta file found!\n"; return 1; }
        std::cout << files.size() << " fasta files have been loaded (run with --verbose to list the files):\n";
        if (a.has("verbose")) for (auto const& file : files) std::cout << file.first << '\n';
    }
    if (ids_lines.empty()) { rmdir(index_dir.c_str()); std::cerr << "ERROR: There is no non-empty sequence in the fasta file(s).\n"; return 1; }

    const uint32_t n_seq = (uint32_t)(limits.size() - 1);
    uint32_t flags = a.has("no-sa") ? 0u : GMB_BUILD_WITH_SA;
    const bool gpu = !a.has("host-build") && gmb_device_count() > 0;
    if (gpu) flags |= GMB_BUILD_ON_GPU;
    if (a.has("verbose")) std::cout << "Building the bidirectional FM index of " << codes.size() << " bases in " << n_seq
                                    << " sequences " << (gpu ? "on the GPU" : "on the host (SA-IS)") << " ... " << std::flush;
    const double t0 = wall();
    void* blob = nullptr;
    uint64_t bytes = 0;
    if (gmb_index_build(codes.data(), limits.data(), n_seq, flags, 0, &blob, &bytes) != GMB_OK) {
        std::cerr << "ERROR: " << gmb_last_error() << "\n";
        rmdir(index_dir.c_str());
        return 1;
    }
    std::string base = index_dir;
    if (base.back() != '/') base += '/';
    if (gmb_blob_save(blob, bytes, (base + "index.gmb").c_str()) != GMB_OK) { std::cerr << "ERROR: " << gmb_last_error() << "\n"; return 1; }
    if (a.has("reference-format")) { // the same index as the fibres the original genmap opens (src/genmap_helper.hpp:71-127)
        uint64_t sampling = 10;
        if (a.has("sampling")) to_uint(a.val["sampling"], sampling);
        std::vector<const char*> id_ptrs;
        for (const std::string& l : ids_lines) id_ptrs.push_back(l.c_str());
        if (gmb_blob_export_reference(blob, bytes, index_dir.c_str(), id_ptrs.data(), (uint32_t)id_ptrs.size(), fd ? 1 : 0, (uint32_t)sampling) != GMB_OK) {
            std::cerr << "ERROR: " << gmb_last_error() << "\n";
            return 1;
        }
    }
    gmb_blob_free(blob);
    {
        std::ofstream ids(base + "index.ids");
        for (const std::string& l : ids_lines) ids << l << '\n';
        std::ofstream info(base + "index.info"); // same keys as src/indexing.hpp:105-111 where they apply
        const bool dna5 = std::find(codes.begin(), codes.end(), (uint8_t)4) != codes.end(); // src/indexing.hpp:459-473
        info << "alphabet_size:" << (dna5 ? 5 : 4) << "\n" << "fasta_directory:" << (fd ? "true" : "false") << "\n"
             << "full_suffix_array:" << (a.has("no-sa") ? "false" : "true") << "\n" << "format:genmap-b200-2\n";
    }
    if (a.has("verbose")) std::cout << "done in " << round2(wall() - t0) << " seconds\n";
    std::cout << "Index created successfully.\n";
    return 0;
}

// ---- map -----------------------------------------------------------------------------------------------
const char* kMapHelp =
    "GenMap map (B200 build)\n\n"
    "    genmap map -I index_dir -O output -K length [-E errors] [-S bed] [-nc] [-ep] [-fs|-fl] -r|-t|-w|-bg|-d\n\n"
    "  -I, --index   -O, --output   -K, --length   -E, --errors (0..4)   -S, --selection\n"
    "  -nc, --no-reverse-complement   -ep, --exclude-pseudo   -fs, --frequency-small   -fl, --frequency-large\n"
    "  -r, --raw   -t, --txt   -w, --wig   -bg, --bedgraph   -d, --csv   -m, --memory-mapping (ignored)\n"
    "  -T, --threads (host threads formatting the txt output; default: all)   -v, --verbose\n"
    "  -xg, --gpus N   range-partition the positions of every FASTA file over N GPUs (index replicated; default 1)\n";

struct IdRow { std::string file; uint64_t length; std::string name; };

// rows of index.ids; `path` is our text file or, for an index written by the reference itself, the SeqAn
// string set index.ids.concat + index.ids.limits (uint64 offsets)
bool load_ids(const std::string& path, std::vector<IdRow>& rows)
{
    std::vector<std::string> lines;
    std::ifstream in(path);
    if (in) {
        std::string l;
        while (std::getline(in, l)) lines.push_back(l);
    } else {
        std::ifstream concat(path + ".concat", std::ios::binary), limits(path + ".limits", std::ios::binary);
        if (!concat || !limits) return false;
        std::string all((std::istreambuf_iterator<char>(concat)), std::istreambuf_iterator<char>());
        std::vector<uint64_t> lim;
        uint64_t v;
        while (limits.read(reinterpret_cast<char*>(&v), 8)) lim.push_back(v);
        for (size_t i = 0; i + 1 < lim.size(); ++i)
            if (lim[i + 1] <= all.size() && lim[i] <= lim[i + 1]) lines.push_back(all.substr(lim[i], lim[i + 1] - lim[i]));
    }
    for (const std::string& line : lines) {
        if (line.empty()) continue;
        const size_t s1 = line.find(';'), s2 = line.find(';', s1 + 1); // src/common.hpp:10-19
        if (s1 == std::string::npos || s2 == std::string::npos) return false;
        rows.push_back(IdRow{line.substr(0, s1), std::stoull(line.substr(s1 + 1, s2 - s1 - 1)), line.substr(s2 + 1)});
    }
    return !rows.empty();
}

// ---- csv (src/output.hpp:189-288) ----------------------------------------------------------------------
// One line per k-mer that has at least one occurrence and whose window stays inside its sequence
// (src/algo.hpp:377-385): "seq,pos" of the k-mer, then per indexed FASTA file the "|"-separated occurrences
// on the + strand, then (unless -nc) the same for the - strand; sequence numbers are local to each file.
// The lists come from gmb_map_locations piece by piece, so memory stays bounded however long the file is
// (the reference keeps a std::map of all k-mers of the file, src/mappability.hpp:168-170).
struct CsvFile { std::string name; uint32_t last_seq; };

bool write_csv(gmb_index* ix, const gmb_params& p, uint64_t text_begin, uint64_t text_len, const std::vector<uint64_t>& cum,
               const std::vector<std::pair<uint64_t, uint64_t>>& iv, const std::vector<CsvFile>& files, const std::string& path)
{
    gmbcli::Sink out(path);
    if (!out.ok()) { std::cerr << "ERROR: cannot write " << path << "\n"; return false; }
    out.str("\"k-mer\"");
    for (const CsvFile& f : files) out.str(";\"+ strand " + f.name + "\"");
    if (p.revcompl)
        for (const CsvFile& f : files) out.str(";\"- strand " + f.name + "\"");
    out.ch('\n');
    size_t chrom = 0;
    for (uint64_t b = 0; b < text_len;) {
        gmb_locations L;
        if (gmb_map_locations(ix, &p, text_begin, text_len, cum.data(), (uint32_t)cum.size() - 1,
                              reinterpret_cast<const uint64_t (*)[2]>(iv.data()), iv.size(), b, text_len, 0, &L) != GMB_OK) {
            std::cerr << "ERROR: " << gmb_last_error() << "\n";
            return false;
        }
        for (uint64_t j = L.pos_begin; j < L.pos_end; ++j) {
            const uint64_t* o = L.offsets + 2 * (j - L.pos_begin);
            if (o[2] == o[0]) continue; // no hit on either strand (k-mers with N; positions that were not searched)
            while (cum[chrom + 1] <= j) ++chrom;
            out.u64(chrom); out.ch(','); out.u64(j - cum[chrom]);
            for (int strand = 0; strand < (p.revcompl ? 2 : 1); ++strand) {
                uint64_t i = o[strand];
                const uint64_t end = o[strand + 1];
                uint32_t before = 0; // sequences in the previous FASTA files
                for (const CsvFile& f : files) {
                    out.ch(';');
                    bool first = true;
                    while (i < end && L.loc[i].seq <= f.last_seq) {
                        if (!first) out.ch('|');
                        out.u64(L.loc[i].seq - before); out.ch(','); out.u64(L.loc[i].pos);
                        first = false;
                        ++i;
                    }
                    before = f.last_seq + 1;
                }
            }
            out.ch('\n');
        }
        b = L.pos_end;
        gmb_locations_free(&L);
    }
    return true;
}

int map_main(int argc, char const** argv)
{
    std::vector<OptSpec> specs = {{"I", "index", true}, {"O", "output", true}, {"E", "errors", true}, {"K", "length", true},
                                  {"S", "selection", true}, {"nc", "no-reverse-complement", false}, {"ep", "exclude-pseudo", false},
                                  {"fs", "frequency-small", false}, {"fl", "frequency-large", false}, {"r", "raw", false},
                                  {"t", "txt", false}, {"w", "wig", false}, {"bg", "bedgraph", false}, {"b", "bed", false},
                                  {"d", "csv", false}, {"m", "memory-mapping", false}, {"T", "threads", true},
                                  {"v", "verbose", false}, {"xo", "overlap", true}, {"xg", "gpus", true},
                                  {"xv", "host-runs", false}};
    Args a;
    int rc = parse_args("GenMap map", specs, argc, argv, a, kMapHelp);
    if (rc == 2) return 0;
    if (rc) return 1;
    for (const char* req : {"index", "output", "length"})
        if (!a.has(req)) {
            const char* s = !strcmp(req, "index") ? "I" : !strcmp(req, "output") ? "O" : "K";
            std::cerr << "GenMap map: Missing value for option: -" << s << ", --" << req << "\n";
            return 1;
        }
    uint64_t K = 0, E = 0, T = 0, xo = 0, gpu = 0;
    if (!to_uint(a.val["length"], K)) { std::cerr << "GenMap map: the given value '" << a.val["length"] << "' cannot be casted to integer\n"; return 1; }
    if (a.has("errors") && !to_uint(a.val["errors"], E)) { std::cerr << "GenMap map: the given value '" << a.val["errors"] << "' cannot be casted to integer\n"; return 1; }
    if (a.has("threads") && !to_uint(a.val["threads"], T)) { std::cerr << "GenMap map: the given value '" << a.val["threads"] << "' cannot be casted to integer\n"; return 1; }
    if (a.has("gpus") && (!to_uint(a.val["gpus"], gpu) || gpu == 0)) { std::cerr << "GenMap map: --gpus needs a positive integer\n"; return 1; }
    if (!a.has("gpus")) gpu = 1;
    const bool raw = a.has("raw"), txt = a.has("txt"), wig = a.has("wig"), bg = a.has("bedgraph"), bed = a.has("bed"), csv = a.has("csv");
    if (!wig && !bg && !bed && !raw && !txt && !csv) {
        std::cerr << "ERROR: Please choose at least one output format (i.e., --wig, --bedgraph, --bed, --raw, --txt, --csv).\n";
        return 1;
    }
    if (a.has("frequency-small") && a.has("frequency-large")) {
        std::cerr << "ERROR: Cannot use both --frequency-small and --frequency-large. Please choose one.\n";
        return 1;
    }
    const gmbcli::OutputType otype = a.has("frequency-small") ? gmbcli::OutputType::frequency_small
                                   : a.has("frequency-large") ? gmbcli::OutputType::frequency_large : gmbcli::OutputType::mappability;
    if (a.has("overlap")) { // src/mappability.hpp:527-541: xo + 1 adjacent k-mers are searched through their common infix
        if (!to_uint(a.val["overlap"], xo)) { std::cerr << "GenMap map: the given value '" << a.val["overlap"] << "' cannot be casted to integer\n"; return 1; }
        const uint64_t mo = std::min<uint64_t>(K - 1, K - E - 2);
        if (xo > mo) { std::cerr << "ERROR: overlap cannot be larger than min(K - 1, K - E - 2) = " << mo << ".\n"; return 1; }
    }
    if (E > 4) { std::cerr << "E > 4 not yet supported.\n"; return 1; } // src/mappability.hpp:187
    if (K < E + 2) { std::cerr << "ERROR: K must be at least E + 2.\n"; return 1; }

    std::string index_dir = a.val["index"];
    if (index_dir.back() != '/') index_dir += '/';
    std::vector<IdRow> rows;
    std::string info_line, info;
    {
        std::ifstream in(index_dir + "index.info");
        if (!in) in.open(index_dir + "index.info.concat"); // an index written by the reference itself
        if (!in) { std::cerr << "ERROR: cannot open the index at " << index_dir << " (index.info missing)\n"; return 1; }
        while (std::getline(in, info_line)) info += info_line + "\n";
    }
    if (!load_ids(index_dir + "index.ids", rows)) { std::cerr << "ERROR: Malformed index.ids file!\n"; return 1; }
    const bool directory = info.find("fasta_directory:true") != std::string::npos;

    // output path (src/mappability.hpp:562-619)
    std::string out_path = a.val["output"];
    bool includes_filename = false;
    if (is_dir(out_path)) {
        if (out_path.back() != '/') out_path += '/';
    } else if (!directory) {
        if (out_path.back() == '.') {
            out_path += '/';
        } else {
            const size_t sl = out_path.find_last_of('/');
            const std::string parent = sl == std::string::npos ? "." : out_path.substr(0, sl);
            includes_filename = true;
            if (!is_dir(parent)) {
                std::cerr << "ERROR: The output cannot be written to the file " << out_path << ".\n"
                          << "       It seems the directory " << parent << " does not exist.\n";
                return 1;
            }
        }
    } else {
        std::cerr << "ERROR: The output directory " << out_path << " does not exist.\n"
                  << "       A filename can only be specified for single indexed fasta files (not for indexed fasta directories).\n"
                  << "       Please create it, or choose a different location.\n";
        return 1;
    }

    // selection (src/mappability.hpp:253-269)
    std::map<std::string, std::vector<std::pair<uint64_t, uint64_t>>> selection;
    const bool has_selection = a.has("selection");
    if (has_selection) {
        std::ifstream in(a.val["selection"]);
        if (!in) { std::cerr << "ERROR: cannot open the bed file " << a.val["selection"] << "\n"; return 1; }
        std::string line;
        while (std::getline(in, line)) {
            std::istringstream ss(line);
            std::string name;
            uint64_t b, e;
            if (line.empty() || line[0] == '#' || !(ss >> name >> b >> e)) continue;
            selection[name].push_back({b, e});
        }
    }

    const double t_cuda = wall();
    if (gmb_device_count() == 0) { std::cerr << "ERROR: no CUDA device found: the B200 build of `genmap map` has no CPU fallback.\n"; return 1; }
    if (a.has("verbose")) std::cout << "CUDA driver initialised in " << round2(wall() - t_cuda) << " seconds\n";
    if ((int)gpu > gmb_device_count()) { std::cerr << "ERROR: --gpus " << gpu << " requested but only " << gmb_device_count() << " CUDA device(s) found.\n"; return 1; }
    std::vector<gmb_index*> ixs(gpu, nullptr); // the index is replicated: one copy in the HBM of every GPU
    const double t_open = wall();
    if (gmb_index_open(index_dir.c_str(), 0, &ixs[0]) != GMB_OK) { std::cerr << "ERROR: " << gmb_last_error() << "\n"; return 1; }
    const double t_repl = wall();
    { // read once, then GPU-to-GPU copies over NVLink, all at once (a CUDA context per device is created on the way)
        std::vector<std::string> rep_err(gpu);
        std::vector<std::thread> rep;
        for (uint64_t g = 1; g < gpu; ++g)
            rep.emplace_back([&, g] { if (gmb_index_replicate(ixs[0], (int)g, &ixs[g]) != GMB_OK) rep_err[g] = gmb_last_error(); });
        for (std::thread& t : rep) t.join();
        for (const std::string& e : rep_err)
            if (!e.empty()) { std::cerr << "ERROR: " << e << "\n"; return 1; }
    }
    const double t_loaded = wall();
    gmb_index_info iinfo;
    gmb_index_get_info(ixs[0], &iinfo);
    if (a.has("verbose")) {
        std::cout << "Index was loaded (" << (iinfo.alphabet_size == 5 ? "dna5" : "dna4") << " alphabet, " << iinfo.blob_bytes << " bytes in the HBM of " << gpu << " GPU(s)).\n";
        std::cout << "- Index read and copied to GPU 0 in " << round2(t_repl - t_open) << " seconds";
        if (gpu > 1) std::cout << ", replicated to " << (gpu - 1) << " more GPU(s) in " << round2(t_loaded - t_repl) << " seconds";
        std::cout << "\n";
        std::cout << (directory ? "- Index was built on an entire directory.\n" : "- Index was built on a single fasta file.\n") << std::flush;
    }

    // file ids per sequence (src/mappability.hpp:230-250)
    std::vector<uint32_t> seq_to_file(rows.size());
    uint32_t total_files = 0;
    for (size_t i = 0; i < rows.size(); ++i) {
        if (i && rows[i].file != rows[i - 1].file) ++total_files;
        seq_to_file[i] = total_files;
    }
    ++total_files;
    std::vector<CsvFile> csv_files; // every indexed FASTA file with its last sequence number (src/output.hpp:200-213)
    for (size_t i = 0; i < rows.size(); ++i)
        if (i + 1 == rows.size() || rows[i + 1].file != rows[i].file) csv_files.push_back(CsvFile{rows[i].file, (uint32_t)i});
    if (csv && !iinfo.has_sa) { std::cerr << "ERROR: --csv needs an index that stores the suffix array (built without --no-sa).\n"; return 1; }
    const bool want_freq = raw || txt || wig || bg || bed;
    // only track formats asked for: the runs are found on the GPU and the vector never comes to the host
    // (--host-runs keeps the host scan, for comparison)
    const bool device_runs = want_freq && !raw && !txt && !a.has("host-runs");

    gmb_params p{};
    p.K = (uint32_t)K; p.E = (uint32_t)E;
    p.revcompl = !a.has("no-reverse-complement");
    p.exclude_pseudo = a.has("exclude-pseudo");
    p.value_bits = otype == gmbcli::OutputType::frequency_small ? 8 : 16; // floats derive from uint16 (:390-393)
    p.block_kmers = a.has("overlap") ? (uint32_t)xo + 1u : 0u;            // stepSize = K - overlap + 1 (src/algo.hpp:416); 0 = default

    const double t_start = wall();
    uint64_t start_pos = 0;
    uint32_t file_no = 0;
    for (size_t i = 0; i < rows.size();) {
        size_t j = i;
        while (j < rows.size() && rows[j].file == rows[i].file) ++j;
        ++file_no;
        std::vector<std::string> names;
        std::vector<uint64_t> lens, cum{0};
        std::vector<std::pair<uint64_t, uint64_t>> iv;
        for (size_t r = i; r < j; ++r) {
            auto it = selection.find(rows[r].name);
            if (it != selection.end())
                for (auto const& x : it->second) {
                    if (x.first >= rows[r].length || x.second > rows[r].length) { // :343-349
                        std::cerr << "Error in BED file! Coordinates exceed sequence length: Seq. \"" << rows[r].name
                                  << "\" has a length of " << rows[r].length << ", but half-closed interval [" << x.first
                                  << ", " << x.second << ") given.\n";
                        return 1;
                    }
                    iv.push_back({cum.back() + x.first, cum.back() + x.second});
                }
            names.push_back(rows[r].name);
            lens.push_back(rows[r].length);
            cum.push_back(cum.back() + rows[r].length);
        }
        const uint64_t text_len = cum.back();
        if (!(has_selection && iv.empty())) { // :309 — files without selected intervals produce no output
            // the frequency vector of the file in host memory: every byte is overwritten by the device-to-host copies, so
            // it is not zero-filled; its pages are touched by several threads first (6 GB at 3 Gbp: a value-initialised
            // std::vector spent 2 s in page faults on one thread)
            const size_t c_bytes = want_freq && !device_runs ? text_len * (p.value_bits / 8) : 0;
            std::unique_ptr<uint8_t[]> c_mem(c_bytes ? new uint8_t[c_bytes] : nullptr);
            struct { uint8_t* p; uint8_t* data() const { return p; } } c{c_mem.get()};
            if (c_bytes) {
                const unsigned nt = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
                std::vector<std::thread> touch;
                for (unsigned t = 0; t < nt; ++t)
                    touch.emplace_back([&, t] {
                        const size_t b = c_bytes * t / nt, e = c_bytes * (t + 1) / nt;
                        for (size_t i = b; i < e; i += 4096) c_mem[i] = 0;
                    });
                for (std::thread& w : touch) w.join();
            }
            std::vector<uint64_t> run_start; // device_runs: the slices of all GPUs, concatenated
            std::vector<uint16_t> run_value;
            std::vector<gmb_runs> slices(gpu);
            static_assert(sizeof(std::pair<uint64_t, uint64_t>) == 16, "interval layout");
            // positions are range-partitioned over the GPUs; every GPU fills its own slice of c
            std::vector<std::string> errors(gpu);
            std::vector<std::thread> workers;
            const double t_search = wall();
            // progress of the call in flight, as the reference prints it (src/common.hpp:94-131): polled from the devices
            std::atomic<bool> searching{true};
            std::thread progress([&] {
                const double t0 = wall();
                while (searching.load()) {
                    std::this_thread::sleep_for(std::chrono::milliseconds(100));
                    if (!searching.load() || wall() - t0 < 0.5) continue;
                    uint64_t done = 0, total = 0;
                    for (uint64_t g = 0; g < gpu; ++g) {
                        uint64_t d = 0, t = 0;
                        if (gmb_progress(ixs[g], &d, &t) == GMB_OK) { done += d; total += t; }
                    }
                    if (total == 0) continue;
                    char buf[64];
                    std::snprintf(buf, sizeof buf, "%.2f", 100.0 * (double)done / (double)total);
                    if (total_files == 1) std::cout << "\rProgress: " << buf << "%\x1b[K" << std::flush;
                    else std::cout << "\rFile " << file_no << " / " << total_files << ". Progress: " << buf << " %\x1b[K" << std::flush;
                }
            });
            for (uint64_t g = 0; want_freq && g < gpu; ++g)
                workers.emplace_back([&, g] {
                    const uint64_t b = text_len * g / gpu, e = text_len * (g + 1) / gpu;
                    if (device_runs) {
                        if (gmb_map_runs(ixs[g], &p, start_pos, text_len, cum.data(), (uint32_t)lens.size(),
                                         reinterpret_cast<const uint64_t (*)[2]>(iv.data()), iv.size(), seq_to_file.data(),
                                         (uint32_t)seq_to_file.size(), b, e, &slices[g], nullptr) != GMB_OK)
                            errors[g] = gmb_last_error();
                    } else if (gmb_map_frequencies_range(ixs[g], &p, start_pos, text_len, cum.data(), (uint32_t)lens.size(),
                                                  reinterpret_cast<const uint64_t (*)[2]>(iv.data()), iv.size(), seq_to_file.data(),
                                                  (uint32_t)seq_to_file.size(), b, e, c.data() + b * (p.value_bits / 8), nullptr) != GMB_OK)
                        errors[g] = gmb_last_error();
                });
            for (std::thread& w : workers) w.join();
            searching.store(false);
            progress.join();
            const double t_searched = wall();
            for (const std::string& e : errors)
                if (!e.empty()) { std::cerr << "ERROR: " << e << "\n"; return 1; }
            if (device_runs) { // a run that continues across a slice boundary is one run
                for (uint64_t g = 0; g < gpu; ++g) {
                    for (uint64_t r = 0; r < slices[g].n_runs; ++r) {
                        const uint64_t st = slices[g].start[r];
                        const bool seq_start = std::binary_search(cum.begin(), cum.end(), st);
                        if (r == 0 && !run_start.empty() && !seq_start && run_value.back() == slices[g].value[r]) continue;
                        run_start.push_back(st);
                        run_value.push_back(slices[g].value[r]);
                    }
                    gmb_runs_free(&slices[g]);
                }
            }
            if (total_files == 1) std::cout << "\rProgress: 100.00%\x1b[K\n" << std::flush;
            else {
                std::cout << "\rFile " << file_no << " / " << total_files << ". Progress: 100.00 %\x1b[K" << std::flush;
                if (a.has("verbose") || file_no == total_files) std::cout << '\n';
            }
            if (a.has("verbose") && want_freq)
                std::cout << "- Searched on " << gpu << " GPU(s) in " << round2(t_searched - t_search) << " seconds ("
                          << (device_runs ? "runs" : "frequency vector") << " in host memory)\n";
            std::string prefix = out_path;
            if (!includes_filename) prefix += rows[i].file.substr(0, rows[i].file.find_last_of('.')) + ".genmap"; // :76-78
            gmbcli::Outputs o{raw, txt, wig, bg, bed, a.has("verbose"),
                              (unsigned)(a.has("threads") ? std::max<uint64_t>(1, T) : std::max(1u, std::thread::hardware_concurrency()))};
            if (!want_freq) {}
            else if (device_runs)
                gmbcli::write_track_outputs(gmbcli::ListRuns{run_start.data(), run_value.data(), run_start.size(), cum}, prefix, names, lens, otype, o);
            else if (p.value_bits == 8) gmbcli::write_outputs(c.data(), text_len, prefix, names, lens, otype, o);
            else gmbcli::write_outputs(reinterpret_cast<const uint16_t*>(c.data()), text_len, prefix, names, lens, otype, o);
            if (csv) {
                const double t_csv = wall();
                if (!write_csv(ixs[0], p, start_pos, text_len, cum, iv, csv_files, prefix + ".csv")) return 1;
                if (a.has("verbose")) std::cout << "- CSV file written in " << round2(wall() - t_csv) << " seconds\n";
            }
        }
        start_pos += text_len;
        i = j;
    }
    if (a.has("verbose")) std::cout << "Mappability computed in " << round2(wall() - t_start) << " seconds\n";
    // every output file is closed; the index replicas and their tables (tens of GB per GPU) are not freed one by one:
    // the process ends here and the driver reclaims them
    std::cout << std::flush;
    std::cerr << std::flush;
    std::fflush(nullptr);
    std::_Exit(0);
}

// hidden developer command: render the text/track formats from a raw frequency file (lets the CPU-only
// test-suite check the writers against the reference's golden outputs without a GPU)
int render_main(int argc, char const** argv)
{
    std::vector<OptSpec> specs = {{"I", "ids", true}, {"C", "counts", true}, {"O", "output", true}, {"N", "file-no", true},
                                  {"fs", "frequency-small", false}, {"fl", "frequency-large", false}, {"r", "raw", false},
                                  {"t", "txt", false}, {"w", "wig", false}, {"bg", "bedgraph", false}, {"b", "bed", false},
                                  {"xr", "via-runs", false}, {"T", "threads", true}};
    Args a;
    int rc = parse_args("GenMap render", specs, argc, argv, a, "genmap render -I index.ids -C counts.freq16|.freq8 -N file_no -O prefix [-fs|-fl] -r -t -w -bg -b\n");
    if (rc) return rc == 2 ? 0 : 1;
    std::vector<IdRow> rows;
    if (!load_ids(a.val["ids"], rows)) { std::cerr << "ERROR: Malformed index.ids file!\n"; return 1; }
    uint64_t file_no = 0;
    if (a.has("file-no")) to_uint(a.val["file-no"], file_no);
    std::vector<std::string> names;
    std::vector<uint64_t> lens;
    uint64_t f = 0;
    for (size_t i = 0; i < rows.size(); ++i) {
        if (i && rows[i].file != rows[i - 1].file) ++f;
        if (f == file_no) { names.push_back(rows[i].name); lens.push_back(rows[i].length); }
    }
    uint64_t n = 0;
    for (uint64_t l : lens) n += l;
    const std::string cpath = a.val["counts"];
    const bool in8 = cpath.size() > 6 && cpath.substr(cpath.size() - 6) == ".freq8";
    std::ifstream in(cpath, std::ios::binary);
    std::vector<uint8_t> buf(n * (in8 ? 1 : 2));
    if (!in.read(reinterpret_cast<char*>(buf.data()), (std::streamsize)buf.size())) { std::cerr << "ERROR: short counts file\n"; return 1; }
    const gmbcli::OutputType otype = a.has("frequency-small") ? gmbcli::OutputType::frequency_small
                                   : a.has("frequency-large") ? gmbcli::OutputType::frequency_large : gmbcli::OutputType::mappability;
    uint64_t n_threads = 3;
    if (a.has("threads")) to_uint(a.val["threads"], n_threads);
    gmbcli::Outputs o{a.has("raw"), a.has("txt"), a.has("wig"), a.has("bedgraph"), a.has("bed"), false, (unsigned)std::max<uint64_t>(1, n_threads)};
    if (a.has("via-runs")) { // the track writers fed from a run list shaped like gmb_map_runs' (a run starts at every sequence start)
        const std::vector<uint64_t> cum = gmbcli::cumulative(lens);
        std::vector<uint64_t> st;
        std::vector<uint16_t> val;
        auto at = [&](uint64_t i) { return in8 ? (uint16_t)buf[i] : reinterpret_cast<const uint16_t*>(buf.data())[i]; };
        for (uint64_t i = 0; i < n; ++i)
            if (i == 0 || at(i) != at(i - 1) || std::binary_search(cum.begin(), cum.end(), i)) { st.push_back(i); val.push_back(at(i)); }
        gmbcli::write_track_outputs(gmbcli::ListRuns{st.data(), val.data(), st.size(), cum}, a.val["output"], names, lens, otype, o);
        return 0;
    }
    if (in8) gmbcli::write_outputs(buf.data(), n, a.val["output"], names, lens, otype, o);
    else gmbcli::write_outputs(reinterpret_cast<const uint16_t*>(buf.data()), n, a.val["output"], names, lens, otype, o);
    return 0;
}

const char* kMainHelp =
    "GenMap - Fast and Exact Computation of Genome Mappability (B200 build)\n\n"
    "    genmap [OPTIONS] COMMAND [COMMAND-OPTIONS]\n\n"
    "Available commands\n"
    "    index  - Creates an index for mappability computation.\n"
    "    map    - Computes the mappability (requires a pre-built index).\n"
    "To view the help page for a specific command, simply run 'genmap command --help'.\n";

} // namespace

int main(int argc, char const** argv)
{
    if (argc < 2) { std::cerr << "GenMap: Too few arguments!\n" << kMainHelp; return 1; }
    const std::string cmd = argv[1];
    if (cmd == "--help" || cmd == "-h") { std::cout << kMainHelp; return 0; }
    if (cmd == "--version") { std::cout << "GenMap version: " << kVersion << "\n" << gmb_version() << "\n"; return 0; }
    if (cmd == "--copyright") { std::cout << "genmap-b200; CLI surface after GenMap (3-clause BSD).\n"; return 0; }
    if (cmd == "index") return index_main(argc - 1, argv + 1);
    if (cmd == "map") return map_main(argc - 1, argv + 1);
    if (cmd == "render") return render_main(argc - 1, argv + 1);
    std::cerr << "GenMap: the given value '" << cmd << "' is not in the list of allowed values [index, map]\n";
    return 1;
}
