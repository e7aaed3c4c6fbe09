// bc_host.cpp — host side of the drop-in (include/bc_host.h): scheme/CSV set-up, FASTQ ingest + packing,
// CSV writers.  Index based: barcodes are reference indices or packed raw keys end to end; strings appear
// only when files are read and written.  No read is decoded here — that is the GPU library's job.
#include "../../../include/bc_host.h"

#include <cuda_runtime_api.h>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <sched.h>
#include <zlib.h>
#if defined(__x86_64__)
#include <immintrin.h>
#endif

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <mutex>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <functional>
#include <map>
#include <memory>
#include <set>
#include <sstream>
#include <stdexcept>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

namespace {

struct Error : std::runtime_error {
    using std::runtime_error::runtime_error;
};

std::string slurp(const std::string& path, const char* verb) {
    std::ifstream in(path, std::ios::binary);
    if (!in) throw Error(std::string(verb) + " " + path);
    std::ostringstream ss;
    ss << in.rdbuf();
    return ss.str();
}

// text -> lines without terminators ("\n" and "\r\n"), no phantom last line
std::vector<std::string> text_lines(const std::string& text) {
    std::vector<std::string> lines;
    size_t start = 0;
    while (start < text.size()) {
        size_t end = text.find('\n', start);
        size_t stop = end == std::string::npos ? text.size() : end;
        size_t len = stop - start;
        if (len && text[stop - 1] == '\r') len--;
        lines.emplace_back(text, start, len);
        if (end == std::string::npos) break;
        start = end + 1;
    }
    return lines;
}

std::vector<std::string> csv_fields(const std::string& line) {
    std::vector<std::string> f;
    std::string cur;
    for (char c : line) {
        if (c == ',') {
            f.push_back(cur);
            cur.clear();
        } else {
            cur.push_back(c);
        }
    }
    f.push_back(cur);
    return f;
}

struct SlotInfo {
    char kind;  // 'S' 'B' 'R'
    uint16_t offset, len;
    std::vector<std::string> dna, name;        // reference barcodes in first-seen order (empty: raw)
    std::vector<const char*> dna_ptrs;
    std::unordered_map<std::string, size_t> pos;  // DNA -> index
};

struct ReadRef {
    const char* seq;
    const char* qual;
    uint32_t len, qlen;
};

// ---- FASTQ streaming: big blocks straight from the file (plain) or through zlib's gz layer (gzip; concatenated
// members are walked like flate2's MultiGzDecoder, input.rs:63), records split in place.  A reader thread fills and
// splits block i+1 while the caller packs block i.
struct FastqBlock {
    std::vector<char> buf;
    size_t have = 0;
    std::vector<ReadRef> recs;
    int state = 0;  // 0 free (reader may fill), 1 filled (consumer may pack)
};

// BGZF (bgzip): a gzip file made of independent members of at most 64 KB, each carrying its compressed size in a
// 'BC' extra subfield (SAM specification, section 4.1).  The member boundaries are known without inflating
// anything, so the members that fit the caller's buffer are inflated on all host threads at once, each straight to
// its place in the buffer — the reference's MultiGzDecoder (input.rs:63) inflates them one after the other.
struct BgzfReader {
    const unsigned char* data = nullptr;
    size_t size = 0, pos = 0;
    int fd = -1;
    unsigned threads = 1;
    ~BgzfReader() {
        if (data) munmap(const_cast<unsigned char*>(data), size);
        if (fd >= 0) close(fd);
    }
    static uint32_t le16(const unsigned char* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8); }
    static uint32_t le32(const unsigned char* p) { return le16(p) | (le16(p + 2) << 16); }
    // total size of the member at p (0: not a BGZF member), and where its deflate stream starts
    static size_t member(const unsigned char* p, size_t avail, size_t* cdata_off) {
        if (avail < 18 || p[0] != 0x1f || p[1] != 0x8b || p[2] != 8 || !(p[3] & 4)) return 0;
        const uint32_t xlen = le16(p + 10);
        if (12 + (size_t)xlen > avail) return 0;
        for (uint32_t x = 0; x + 4 <= xlen;) {
            const unsigned char* sf = p + 12 + x;
            const uint32_t slen = le16(sf + 2);
            if (sf[0] == 'B' && sf[1] == 'C' && slen == 2 && x + 6 <= xlen) {
                *cdata_off = 12 + (size_t)xlen;
                return (size_t)le16(sf + 4) + 1;
            }
            x += 4 + slen;
        }
        return 0;
    }
    bool open(const std::string& path, unsigned n_threads) {
        fd = ::open(path.c_str(), O_RDONLY);
        if (fd < 0) return false;
        struct stat st;
        if (fstat(fd, &st) != 0 || !S_ISREG(st.st_mode) || st.st_size < 28) return false;
        void* m = mmap(nullptr, (size_t)st.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
        if (m == MAP_FAILED) return false;
        data = static_cast<const unsigned char*>(m);
        size = (size_t)st.st_size;
        threads = std::max(1u, n_threads);
        size_t off = 0;
        return member(data, size, &off) != 0;
    }
    // whole members, as many as fit n bytes; 0 at the end of the file, -2 when the next member does not fit
    long fill(char* dst, size_t n) {
        struct Blk {
            size_t in, in_len, out;
            uint32_t isize, crc;
        };
        std::vector<Blk> blks;
        size_t out = 0, p = pos;
        while (p < size) {
            size_t coff = 0;
            const size_t total = member(data + p, size - p, &coff);
            if (total == 0 || p + total > size || total < coff + 8) throw Error("corrupt or truncated BGZF member in the input");
            const uint32_t isize = le32(data + p + total - 4);
            if (out + isize > n) {
                if (blks.empty() && out == 0) return -2;  // not even one member fits what is left of the buffer
                break;
            }
            if (isize) blks.push_back(Blk{p + coff, total - coff - 8, out, isize, le32(data + p + total - 8)});
            out += isize;
            p += total;
        }
        std::atomic<size_t> next{0};
        std::atomic<int> bad{0};
        auto work = [&]() {
            z_stream z;
            memset(&z, 0, sizeof z);
            if (inflateInit2(&z, -15) != Z_OK) {
                bad = 1;
                return;
            }
            for (size_t i = next++; i < blks.size(); i = next++) {
                const Blk& b = blks[i];
                inflateReset(&z);
                z.next_in = const_cast<unsigned char*>(data + b.in);
                z.avail_in = (uInt)b.in_len;
                z.next_out = reinterpret_cast<unsigned char*>(dst + b.out);
                z.avail_out = b.isize;
                const int rc = inflate(&z, Z_FINISH);
                if (rc != Z_STREAM_END || z.avail_out != 0 ||
                    (uint32_t)crc32(crc32(0L, Z_NULL, 0), reinterpret_cast<const unsigned char*>(dst + b.out), b.isize) != b.crc)
                    bad = 1;
            }
            inflateEnd(&z);
        };
        const unsigned nt = (unsigned)std::min<size_t>(threads, std::max<size_t>(1, blks.size() / 8));
        if (nt <= 1) {
            work();
        } else {
            std::vector<std::thread> pool;
            for (unsigned t = 0; t < nt; t++) pool.emplace_back(work);
            for (auto& th : pool) th.join();
        }
        if (bad) throw Error("inflate failed on a BGZF member (corrupt input?)");
        pos = p;
        return (long)out;
    }
};

struct FastqStream {
    gzFile gz = nullptr;
    FILE* plain = nullptr;  // not gzip: read straight into the block buffer (no inflate layer, no extra copy)
    std::unique_ptr<BgzfReader> bgzf;  // bgzip input: members inflated in parallel
    bool eof = false;
    std::vector<char> carry;  // bytes after the last whole record of the previous block
    FastqStream(const std::string& path, unsigned threads = 1) {
        auto ends = [&](const char* suf) {
            const size_t k = strlen(suf);
            return path.size() >= k && path.compare(path.size() - k, k, suf) == 0;
        };
        if (!ends("fastq") && !ends("fastq.gz"))  // input.rs:33-39
            throw Error("This program only works with *.fastq files and *.fastq.gz files.  The latter is still experimental");
        FILE* probe = fopen(path.c_str(), "rb");
        if (!probe) throw Error("Failed to open file: " + path);
        unsigned char magic[2] = {0, 0};
        const size_t got = fread(magic, 1, 2, probe);
        if (got == 2 && magic[0] == 0x1f && magic[1] == 0x8b) {
            fclose(probe);
            bgzf.reset(new BgzfReader());
            if (bgzf->open(path, threads)) return;
            bgzf.reset();
            gz = gzopen(path.c_str(), "rb");
            if (!gz) throw Error("Failed to open file: " + path);
            gzbuffer(gz, 4u << 20);
        } else {
            rewind(probe);
            setvbuf(probe, nullptr, _IONBF, 0);
            plain = probe;
        }
    }
    ~FastqStream() {
        if (gz) gzclose(gz);
        if (plain) fclose(plain);
    }
    long fill(char* dst, size_t n) {
        if (bgzf) return bgzf->fill(dst, n);
        if (plain) return (long)fread(dst, 1, n, plain);
        return gzread(gz, dst, (unsigned)std::min<size_t>(n, 1u << 30));
    }
    // Fills `B` with up to max_records whole records (pointers into B.buf).  Returns false when the input is exhausted.
    bool next_block(FastqBlock& B, size_t max_records) {
        B.recs.clear();
        B.have = 0;
        if (eof && carry.empty()) return false;
        if (carry.size() > B.buf.size()) throw Error("FASTQ record longer than the block buffer");
        memcpy(B.buf.data(), carry.data(), carry.size());
        B.have = carry.size();
        carry.clear();
        while (!eof && B.have < B.buf.size()) {
            const long got = fill(B.buf.data() + B.have, B.buf.size() - B.have);
            if (got == -2) {  // BGZF: the next member needs more room than is left in this block
                if (B.have == 0) throw Error("BGZF member larger than the block buffer");
                break;
            }
            if (got < 0) throw Error("gzread failed (corrupt input?)");
            if (got == 0) {
                eof = true;
                break;
            }
            B.have += (size_t)got;
        }
        char* base = B.buf.data();
        size_t pos = 0;
        while (B.recs.size() < max_records) {
            const char* line[4];
            uint32_t len[4];
            size_t p = pos;
            int k = 0;
            for (; k < 4; k++) {
                const char* nl = (const char*)memchr(base + p, '\n', B.have - p);
                size_t end;
                if (nl) end = (size_t)(nl - base);
                else if (eof && k == 3 && p < B.have) end = B.have;  // last line without '\n'
                else break;
                size_t l = end - p;
                if (l && base[end - 1] == '\r') l--;
                // input.rs:133-137: the reference drops the last character of a record as "the newline"; on its gzip path a
                // last line without one loses its last quality character instead (its plain path adds the newline first)
                if (!nl && !plain && l) l--;
                line[k] = base + p;
                len[k] = (uint32_t)l;
                p = nl ? end + 1 : end;
            }
            if (k < 4) break;
            B.recs.push_back(ReadRef{line[1], line[3], len[1], len[3]});
            pos = p;
        }
        if (B.recs.empty()) {
            if (eof) return false;  // trailing partial record (fewer than 4 lines) is dropped, as the reference never posts it
            throw Error("FASTQ record longer than the block buffer");
        }
        carry.assign(base + pos, base + B.have);
        return true;
    }
};

struct PinnedBatch {
    uint32_t* planes = nullptr;
    uint16_t* read_len = nullptr;
    uint8_t* qual = nullptr;
    size_t cap_planes = 0, cap_qual = 0;  // bytes
    unsigned char* arena = nullptr;       // transfer form (bc_wire_batch): one allocation, laid out by WireLayout
    size_t cap_arena = 0;
    ~PinnedBatch() { release(); }
    void release() {
        if (planes) cudaFreeHost(planes);
        if (read_len) cudaFreeHost(read_len);
        if (qual) cudaFreeHost(qual);
        if (arena) cudaFreeHost(arena);
        planes = nullptr;
        read_len = nullptr;
        qual = nullptr;
        arena = nullptr;
        cap_arena = 0;
    }
    void alloc_wire(size_t bytes) {
        release();
        if (cudaHostAlloc((void**)&arena, bytes, cudaHostAllocDefault) != cudaSuccess) throw Error("cudaHostAlloc failed for the pinned batch buffers");
        cap_arena = bytes;
    }
    void alloc(uint32_t n, uint32_t max_read_len, bool with_qual) {
        release();
        cap_planes = (size_t)n * bc_plane_stride(max_read_len) * 4;
        cap_qual = with_qual ? (size_t)n * bc_qual_stride(max_read_len) : 0;
        if (cudaHostAlloc((void**)&planes, cap_planes, cudaHostAllocDefault) != cudaSuccess ||
            cudaHostAlloc((void**)&read_len, (size_t)n * 2, cudaHostAllocDefault) != cudaSuccess ||
            (with_qual && cudaHostAlloc((void**)&qual, cap_qual, cudaHostAllocDefault) != cudaSuccess))
            throw Error("cudaHostAlloc failed for the pinned batch buffers");
    }
};

// one GPU of an ingest: two pinned batches that alternate (the copy of one overlaps the packing of the other)
struct Lane {
    PinnedBatch pinned[2];
    int cur = 0, in_flight = 0;
    int device = -1;
};

class Pool;

// ingest state, kept by the run so that repeated bch_count_fastq calls pay neither the page-locking nor the thread
// creation again
struct IngestBuffers {
    std::vector<std::unique_ptr<Lane>> lanes;
    FastqBlock blocks[2];
    std::shared_ptr<Pool> pool;
    uint32_t batch_reads = 0, mrl = 0;
    bool with_qual = false, wire = false;
};

struct IngestStats {  // wall time of the phases of the last ingest, on the ingest thread (bch_ingest_stats)
    double split_s = 0, pack_s = 0, submit_s = 0, wait_s = 0, total_s = 0, map_s = 0;
    uint64_t batches = 0, batches_qual8 = 0, batches_dense_n = 0, h2d_bytes = 0;
};

}  // namespace

struct bch_run {
    std::string format_string, regions_string;
    uint16_t constant_len = 0;
    std::vector<SlotInfo> slots;  // template order
    std::vector<int> counted;     // slot indices of {n} barcodes, in order
    int sample_slot = -1, random_slot = -1;
    bool have_sample_file = false, have_counted_file = false;
    uint16_t max_constant = 0, max_sample = 0;
    std::vector<uint16_t> max_counted, counted_sizes;
    float min_quality = 0.f;
    bc_config cfg{};
    std::string description;
    IngestBuffers ingest;
    bch_progress_fn progress = nullptr;  // called with the running number of records after every batch
    void* progress_user = nullptr;
    uint64_t lean_rows = 4u << 20;  // tables of at least this many rows go through the lean CSV writer (bch_set_option)
    bool wire_batches = true;       // host batches cross PCIe in their transfer form (bc_submit_wire); 0: as bc_batch
    bool fused_ingest = true;       // plain files: frame and pack in one pass over the text (FusedWalk); 0: split, then pack
    int mmap_populate = 0;          // measurement: 1 maps a plain file with MAP_POPULATE, 2 lets every worker populate its chunk
    IngestStats stats;
    int multi_mode = 0;  // after bch_count_fastq_multi: 1 = rows partitioned over the contexts, 2 = all rows on the first
};

namespace {

// ---- scheme file: tokens {n} [n] (n), runs of N, runs of ACGT; everything else is ignored; lines starting with
// '#' are comments and the remaining lines are joined without a separator (info.rs:218-233)
void parse_scheme(bch_run& run, const std::string& text) {
    std::string body;
    for (const std::string& line : text_lines(text))
        if (line.empty() || line[0] != '#') body += line;
    auto digit = [](char c) { return c >= '0' && c <= '9'; };
    size_t i = 0;
    const size_t n = body.size();
    while (i < n) {
        const char c = body[i];
        if (c == '{' || c == '[' || c == '(') {
            const char closer = c == '{' ? '}' : c == '[' ? ']' : ')';
            size_t j = i + 1;
            unsigned long value = 0;
            while (j < n && digit(body[j])) {
                value = value * 10 + (unsigned long)(body[j] - '0');
                if (value > 65535) throw Error("scheme: barcode length does not fit 16 bits");
                j++;
            }
            if (j == i + 1 || j >= n || body[j] != closer) {
                i++;
                continue;
            }
            SlotInfo s;
            s.kind = c == '{' ? 'B' : c == '[' ? 'S' : 'R';
            s.offset = (uint16_t)run.format_string.size();
            s.len = (uint16_t)value;
            const int idx = (int)run.slots.size();
            if (s.kind == 'S') {
                if (run.sample_slot >= 0) throw Error("scheme: more than one sample barcode [n] (the reference's regex rejects a duplicate group too)");
                run.sample_slot = idx;
            } else if (s.kind == 'R') {
                if (run.random_slot >= 0) throw Error("scheme: more than one random barcode (n) (the reference's regex rejects a duplicate group too)");
                run.random_slot = idx;
            } else {
                run.counted.push_back(idx);
                run.counted_sizes.push_back(s.len);
            }
            run.slots.push_back(s);
            run.format_string.append(value, 'N');
            run.regions_string.append(value, s.kind);
            i = j + 1;
        } else if (c == 'N' || c == 'n') {
            size_t j = i;
            while (j < n && (body[j] == 'N' || body[j] == 'n')) {
                if (body[j] == 'n') throw Error("scheme: lower-case 'n' is not supported (upper-case N only)");
                j++;
            }
            run.format_string.append(j - i, 'N');  // no region code for format-N (info.rs:287-295)
            i = j;
        } else if (strchr("ACGTacgt", c)) {
            size_t j = i;
            while (j < n && strchr("ACGTacgt", body[j]) && body[j] != '\0') {
                if (body[j] >= 'a')
                    throw Error("scheme: lower-case constants are not supported: the reference upper-cases them for the regex but "
                                "not for the repair step (info.rs:298-299), which makes such schemes ill-defined");
                j++;
            }
            run.format_string.append(body, i, j - i);
            run.regions_string.append(j - i, 'C');
            run.constant_len = (uint16_t)(run.constant_len + (j - i));
            i = j;
        } else {
            i++;
        }
    }
    if (run.counted.empty()) throw Error("scheme: no counted barcode {n}");
}

// ---- conversion CSVs (info.rs:364-433): header skipped, comma split, no trimming; a later duplicate DNA
// replaces the earlier ID
void add_ref(SlotInfo& s, const std::string& dna, const std::string& id) {
    auto it = s.pos.find(dna);
    if (it != s.pos.end()) {
        s.name[it->second] = id;
        return;
    }
    s.pos.emplace(dna, s.dna.size());
    s.dna.push_back(dna);
    s.name.push_back(id);
}

void load_samples(bch_run& run, const std::string& path) {
    if (run.sample_slot < 0)
        throw Error("--sample-barcodes given but the scheme has no sample barcode [n]: the reference silently drops every count in "
                    "that configuration (info.rs:762-766); refusing to run it");
    std::vector<std::string> lines = text_lines(slurp(path, "Failed to open"));
    SlotInfo& s = run.slots[run.sample_slot];
    for (size_t i = 1; i < lines.size(); i++) {
        std::vector<std::string> f = csv_fields(lines[i]);
        if (f.size() >= 2) add_ref(s, f[0], f[1]);
        else add_ref(s, "", "");
    }
    run.have_sample_file = !s.dna.empty();
}

void load_counted(bch_run& run, const std::string& path) {
    std::vector<std::string> lines = text_lines(slurp(path, "Failed to read"));
    std::vector<bool> seen(run.counted.size(), false);
    for (size_t i = 1; i < lines.size(); i++) {
        std::vector<std::string> f = csv_fields(lines[i]);
        std::string dna, id, num;
        if (f.size() >= 3) {
            dna = f[0];
            id = f[1];
            num = f[2];
        }
        size_t p = (!num.empty() && num[0] == '+') ? 1 : 0;
        bool numeric = p < num.size();
        for (size_t q = p; q < num.size() && numeric; q++) numeric = num[q] >= '0' && num[q] <= '9';
        if (!numeric) throw Error("Third column of barcode file contains something other than an integer: " + num);
        const unsigned long long k = std::stoull(num.substr(p));
        if (k == 0 || k > run.counted.size())
            throw Error("barcode file: barcode number " + num + " is outside 1.." + std::to_string(run.counted.size()));
        seen[k - 1] = true;
        add_ref(run.slots[run.counted[k - 1]], dna, id);
    }
    std::string missing;
    for (size_t k = 0; k < seen.size(); k++)
        if (!seen[k]) missing += (missing.empty() ? "" : ", ") + std::to_string(k);
    if (!missing.empty()) throw Error("Barcode conversion file missing barcode numers [" + missing + "] in the third column");
    run.have_counted_file = true;
}

std::string shortest_float(float v) {
    char buf[64];
    for (int p = 1; p <= 9; p++) {
        snprintf(buf, sizeof buf, "%.*g", p, (double)v);
        if (strtof(buf, nullptr) == v) break;
    }
    return buf;
}

std::string list_u16(const std::vector<uint16_t>& v) {
    std::string s = "[";
    for (size_t i = 0; i < v.size(); i++) s += (i ? ", " : "") + std::to_string(v[i]);
    return s + "]";
}

void describe(bch_run& run) {
    std::string key;
    std::string seen;
    for (char c : run.regions_string) {
        if (seen.find(c) != std::string::npos) continue;
        seen.push_back(c);
        if (c == 'S') key += "\nS: Sample barcode";
        else if (c == 'B') key += "\nB: Counted barcode";
        else if (c == 'C') key += "\nC: Constant region";
        else if (c == 'R') key += "\nR: Random barcode";
    }
    std::string d = "-FORMAT-\n" + run.format_string + "\n" + run.regions_string + key + "\n\n";
    const std::string bar = "--------------------------------------------------------------\n";
    d += "-BARCODE INFO-\nConstant region size: " + std::to_string(run.constant_len) +
         "\nMaximum mismatches allowed per sequence: " + std::to_string(run.max_constant) + "\n" + bar;
    d += "Sample barcode size: " + std::to_string(run.sample_slot >= 0 ? run.slots[run.sample_slot].len : 0) +
         "\nMaximum mismatches allowed per sequence: " + std::to_string(run.max_sample) + "\n" + bar;
    if (run.counted_sizes.size() > 1)
        d += "Barcode sizes: " + list_u16(run.counted_sizes) + "\nMaximum mismatches allowed per barcode sequence: " +
             list_u16(run.max_counted) + "\n";
    else
        d += "Barcode size: " + std::to_string(run.counted_sizes[0]) + "\nMaximum mismatches allowed per barcode sequence: " +
             std::to_string(run.max_counted[0]) + "\n";
    d += bar + "Minimum allowed average read quality score per barcode: " + shortest_float(run.min_quality) + "\n";
    run.description = d;
}

// ---- packing ------------------------------------------------------------------------------------------------

struct BaseLut {
    uint8_t code[256];  // 0..3 = A C G T, 4 = N, 5 = anything else
    BaseLut() {
        memset(code, 5, sizeof code);
        code[(unsigned char)'A'] = 0;
        code[(unsigned char)'C'] = 1;
        code[(unsigned char)'G'] = 2;
        code[(unsigned char)'T'] = 3;
        code[(unsigned char)'N'] = 4;
    }
};
const BaseLut kLut;

// 32 bases -> one word of each plane.  Scalar reference version and an AVX2 version (chosen once at start-up).
inline void pack_word_scalar(const char* seq, uint32_t m, uint32_t* lo, uint32_t* hi, uint32_t* nm, bool* other) {
    uint32_t l = 0, h = 0, x = 0;
    for (uint32_t b = 0; b < m; b++) {
        const uint8_t c = kLut.code[(unsigned char)seq[b]];
        l |= (uint32_t)(c & 1u) << b;
        h |= (uint32_t)((c >> 1) & 1u) << b;
        x |= (uint32_t)(c >> 2) << b;  // N or other
        *other |= c == 5;
    }
    *lo = l & ~x;
    *hi = h & ~x;
    *nm = x;
}

#if defined(__x86_64__)
__attribute__((target("avx2"))) inline void pack_word_avx2(const char* seq, uint32_t* lo, uint32_t* hi, uint32_t* nm, bool* other) {
    const __m256i v = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(seq));
    const uint32_t isA = (uint32_t)_mm256_movemask_epi8(_mm256_cmpeq_epi8(v, _mm256_set1_epi8('A')));
    const uint32_t isC = (uint32_t)_mm256_movemask_epi8(_mm256_cmpeq_epi8(v, _mm256_set1_epi8('C')));
    const uint32_t isG = (uint32_t)_mm256_movemask_epi8(_mm256_cmpeq_epi8(v, _mm256_set1_epi8('G')));
    const uint32_t isT = (uint32_t)_mm256_movemask_epi8(_mm256_cmpeq_epi8(v, _mm256_set1_epi8('T')));
    *lo = isC | isT;
    *hi = isG | isT;
    *nm = ~(isA | isC | isG | isT);  // N or other: the mask convention of the scalar version
    const uint32_t isN = (uint32_t)_mm256_movemask_epi8(_mm256_cmpeq_epi8(v, _mm256_set1_epi8('N')));
    *other |= (*nm & ~isN) != 0;
}
const bool kCpuAvx2 = __builtin_cpu_supports("avx2");
// 64 characters per step, compares straight into mask registers, byte permutes across the whole vector (VBMI)
const bool kCpuAvx512 = __builtin_cpu_supports("avx512bw") && __builtin_cpu_supports("avx512vl") && __builtin_cpu_supports("avx512vbmi");
bool kHaveAvx2 = kCpuAvx2, kHaveAvx512 = kCpuAvx512;  // what the packers use: the CPU's best, unless bch_set_simd_level caps it (tests)
#define BC_AVX512 __attribute__((target("avx512f,avx512bw,avx512vl,avx512vbmi")))
#else
const bool kCpuAvx2 = false, kCpuAvx512 = false;
bool kHaveAvx2 = false, kHaveAvx512 = false;
#endif

// one read -> three bit planes + length word (+ quality bytes)
#if defined(__x86_64__)
// 64 bases per step; a masked load takes the last, partial block without touching a byte beyond the read
BC_AVX512 inline bool pack_planes_avx512(const char* seq, uint32_t len, uint32_t W, uint32_t* lo, uint32_t* hi, uint32_t* nm) {
    const __m512i cA = _mm512_set1_epi8('A'), cC = _mm512_set1_epi8('C'), cG = _mm512_set1_epi8('G'), cT = _mm512_set1_epi8('T'),
                  cN = _mm512_set1_epi8('N');
    uint64_t other = 0;
    uint32_t w = 0;
    for (uint32_t base = 0; base < len; base += 64, w += 2) {
        const uint32_t m = len - base;
        const __mmask64 k = m >= 64 ? ~0ULL : ((1ULL << m) - 1ULL);
        const __m512i v = _mm512_maskz_loadu_epi8(k, seq + base);
        const uint64_t isA = _mm512_mask_cmpeq_epi8_mask(k, v, cA), isC = _mm512_mask_cmpeq_epi8_mask(k, v, cC),
                       isG = _mm512_mask_cmpeq_epi8_mask(k, v, cG), isT = _mm512_mask_cmpeq_epi8_mask(k, v, cT),
                       isN = _mm512_mask_cmpeq_epi8_mask(k, v, cN);
        const uint64_t l = isC | isT, h = isG | isT, x = (uint64_t)k & ~(isA | isC | isG | isT);
        other |= x & ~isN;
        lo[w] = (uint32_t)l;
        hi[w] = (uint32_t)h;
        nm[w] = (uint32_t)x;
        if (w + 1 < W) {
            lo[w + 1] = (uint32_t)(l >> 32);
            hi[w + 1] = (uint32_t)(h >> 32);
            nm[w + 1] = (uint32_t)(x >> 32);
        }
    }
    for (; w < W; w++) lo[w] = hi[w] = nm[w] = 0;
    return other != 0;
}
#endif

// the bases of one read -> W words of each plane; true when the read holds a character outside ACGTN
inline bool pack_planes(const char* seq, uint32_t len, uint32_t W, uint32_t* lo, uint32_t* hi, uint32_t* nm) {
    bool other = false;
    uint32_t w = 0, base = 0;
#if defined(__x86_64__)
    if (kHaveAvx512) return pack_planes_avx512(seq, len, W, lo, hi, nm);
    if (kHaveAvx2) {
        for (; base + 32 <= len; base += 32, w++) pack_word_avx2(seq + base, &lo[w], &hi[w], &nm[w], &other);
        if (base < len) {  // the last, partial word through the same vector code: a padded copy ('A' = 00, never N or other)
            char tail[32];
            const uint32_t m = len - base;
            memset(tail, 'A', sizeof tail);
            memcpy(tail, seq + base, m);
            pack_word_avx2(tail, &lo[w], &hi[w], &nm[w], &other);
            w++;
        }
        for (; w < W; w++) lo[w] = hi[w] = nm[w] = 0;
        return other;
    }
#endif
    for (; base < len; base += 32, w++) pack_word_scalar(seq + base, std::min(32u, len - base), &lo[w], &hi[w], &nm[w], &other);
    for (; w < W; w++) lo[w] = hi[w] = nm[w] = 0;
    return other;
}

inline void pack_one(const char* seq, uint32_t len, const char* qual, uint32_t qlen, uint32_t W, uint32_t* planes, uint16_t* read_len,
                     uint8_t* qual_out, uint32_t qual_stride) {
    bool other = pack_planes(seq, len, W, planes, planes + W, planes + 2 * W);
    if (3 * W != ((3 * W) | 1u)) planes[3 * W] = 0;  // pad word of an even record
    if (qual_out) {
        // a quality character below '!' underflows the reference's `q - 33` (parse.rs:326, Q13): flag it, never decode it
        const uint32_t n = std::min(qlen, len);
        unsigned char lowest = 255;
        for (uint32_t i = 0; i < n; i++) lowest = std::min<unsigned char>(lowest, (unsigned char)qual[i]);
        if (n && lowest < 33) other = true;
        memcpy(qual_out, qual, n);
        if (qlen < len) {
            // A quality line shorter than its sequence: the reference zips scores with region codes (parse.rs:338-343), so the
            // walk ends with the line and a barcode run is tested only if the line goes at least one score beyond it.  A score
            // of 255 from the line's last position on makes the sum of any run that reaches it pass every threshold.
            if (n) qual_out[n - 1] = 0xFF;
            memset(qual_out + n, 0xFF, qual_stride - n);
        } else {
            memset(qual_out + len, '!', qual_stride - len);
        }
    }
    *read_len = (uint16_t)(len | (other ? BC_READ_UNSUPPORTED : 0u));
}

// a read this build cannot represent (longer than BC_MAX_READ_LEN): an empty record flagged unsupported, counted separately
inline void pack_unsupported(uint32_t W, uint32_t* planes, uint16_t* read_len, uint8_t* qual_out, uint32_t qual_stride) {
    memset(planes, 0, (size_t)((3 * W) | 1u) * 4);
    if (qual_out) memset(qual_out, '!', qual_stride);
    *read_len = (uint16_t)BC_READ_UNSUPPORTED;
}

// reads [0, count) of `reads` -> batch arrays of geometry max_read_len.  strict: a read longer than max_read_len is an error
// (the caller chose the geometry); otherwise only reads beyond BC_MAX_READ_LEN can be too long and are flagged unsupported.
bool pack_range(uint32_t max_read_len, const ReadRef* reads, size_t count, uint32_t* planes, uint16_t* read_len, uint8_t* qual,
                bool strict) {
    const uint32_t W = bc_plane_words(max_read_len), ps = bc_plane_stride(max_read_len), qs = bc_qual_stride(max_read_len);
    bool ok = true;
    for (size_t i = 0; i < count; i++) {
        const ReadRef& r = reads[i];
        if (r.len > max_read_len || r.len > 0x7FFF) {
            if (strict) ok = false;
            pack_unsupported(W, planes + i * ps, read_len + i, qual ? qual + i * qs : nullptr, qs);
            continue;
        }
        pack_one(r.seq, r.len, r.qual, r.qlen, W, planes + i * ps, read_len + i, qual ? qual + i * qs : nullptr, qs);
    }
    return ok;
}

// ---- transfer form (bc_wire_batch, include/bc_b200.h): lo / hi planes, N calls as a list, quality as 6-bit codes ----------
// One arena per batch; the dense N plane is always written (the list is derived from it and used when it is the smaller).
inline size_t up256(size_t x) { return (x + 255) & ~(size_t)255; }
struct WireLayout {
    uint32_t W = 0, n_codes = 0, list_cap = 0;
    size_t o_lohi = 0, o_len = 0, o_nm = 0, o_nr = 0, o_np = 0, o_q = 0, total = 0;
    WireLayout() = default;
    WireLayout(size_t n, uint32_t mrl, bool with_qual) {
        W = bc_plane_words(mrl);
        n_codes = bc_wire_qual_codes(mrl);
        list_cap = (uint32_t)std::min<size_t>(n * W * 4 / 6 + 1, 0xFFFFFFF0u);  // beyond this many calls the dense plane is smaller
        o_lohi = 0;
        o_len = o_lohi + up256(n * 2 * W * 4);
        o_nm = o_len + up256(n * 2);
        o_nr = o_nm + up256(n * W * 4);
        o_np = o_nr + up256((size_t)list_cap * 4);
        o_q = o_np + up256((size_t)list_cap * 2);
        total = o_q + (with_qual ? up256(n * bc_wire_qual_stride(mrl, 8)) : 0);
    }
};

// the quality line of one read as the decode kernel wants it (see pack_one), n_codes characters
inline void qual_chars(const char* qual, uint32_t qlen, uint32_t len, uint32_t n_codes, uint8_t* out) {
    const uint32_t n = std::min(qlen, len);
    memcpy(out, qual, n);
    if (qlen < len) {
        if (n) out[n - 1] = 0xFF;
        memset(out + n, 0xFF, n_codes - n);
    } else {
        memset(out + len, '!', n_codes - len);
    }
}

// n_codes characters (a multiple of 4) -> 6-bit codes: character - 33, 63 for the 0xFF mark.  false: a character does not
// fit (below '!' is flagged by the caller; '`' and above need the 8-bit form).
inline bool pack_qual6_scalar(const uint8_t* c, uint32_t n_codes, uint8_t* out) {
    bool ok = true;
    for (uint32_t i = 0; i < n_codes; i += 4) {
        uint32_t v = 0;
        for (uint32_t k = 0; k < 4; k++) {
            const uint8_t ch = c[i + k];
            uint32_t code = ch == 0xFF ? 63u : (uint32_t)(uint8_t)(ch - 33);
            if (ch != 0xFF && code > 62u) {
                ok = false;
                code = 0;
            }
            v |= code << (6 * k);
        }
        out[0] = (uint8_t)v;
        out[1] = (uint8_t)(v >> 8);
        out[2] = (uint8_t)(v >> 16);
        out += 3;
    }
    return ok;
}
#if defined(__x86_64__)
// 32 characters -> 24 bytes of codes (shared by the full groups and the padded last one)
__attribute__((target("avx2"))) inline void qual6_group(const __m256i v, __m256i& bad, __m256i& lowest, uint8_t* out24) {
    const __m256i k33 = _mm256_set1_epi8(33), k63 = _mm256_set1_epi8(63), k62 = _mm256_set1_epi8(62), kff = _mm256_set1_epi8((char)0xFF);
    const __m256i mul1 = _mm256_set1_epi16(0x4001), mul2 = _mm256_set1_epi32(0x10000001);
    const __m256i shuf = _mm256_setr_epi8(0, 1, 2, 4, 5, 6, 8, 9, 10, 12, 13, 14, -1, -1, -1, -1, 0, 1, 2, 4, 5, 6, 8, 9, 10, 12, 13, 14, -1, -1, -1, -1);
    const __m256i mark = _mm256_cmpeq_epi8(v, kff);
    lowest = _mm256_min_epu8(lowest, v);
    __m256i code = _mm256_sub_epi8(v, k33);
    bad = _mm256_or_si256(bad, _mm256_andnot_si256(mark, _mm256_xor_si256(_mm256_cmpeq_epi8(_mm256_max_epu8(code, k62), k62), kff)));
    code = _mm256_and_si256(_mm256_blendv_epi8(code, k63, mark), k63);
    const __m256i m = _mm256_shuffle_epi8(_mm256_madd_epi16(_mm256_maddubs_epi16(code, mul1), mul2), shuf);
    const __m128i a = _mm256_castsi256_si128(m), b = _mm256_extracti128_si256(m, 1);
    _mm_storel_epi64(reinterpret_cast<__m128i*>(out24), a);
    const uint32_t a2 = (uint32_t)_mm_extract_epi32(a, 2), b2 = (uint32_t)_mm_extract_epi32(b, 2);
    memcpy(out24 + 8, &a2, 4);
    _mm_storel_epi64(reinterpret_cast<__m128i*>(out24 + 12), b);
    memcpy(out24 + 20, &b2, 4);
}
// the first n_valid characters come from c, the rest of the n_codes (a multiple of 4) are '!'; *lowest_out = smallest character seen
__attribute__((target("avx2"))) inline bool pack_qual6_avx2(const uint8_t* c, uint32_t n_valid, uint32_t n_codes, uint8_t* out, uint8_t* lowest_out) {
    __m256i bad = _mm256_setzero_si256(), lowest = _mm256_set1_epi8((char)0xFF);
    uint32_t i = 0;
    for (; i + 32 <= n_valid; i += 32, out += 24) qual6_group(_mm256_loadu_si256(reinterpret_cast<const __m256i*>(c + i)), bad, lowest, out);
    for (; i < n_codes; i += 32, out += 24) {  // padded groups ('!' = code 0), stored as far as the record goes
        uint8_t tail[32], packed[24];
        memset(tail, '!', sizeof tail);
        if (i < n_valid) memcpy(tail, c + i, n_valid - i);
        qual6_group(_mm256_loadu_si256(reinterpret_cast<const __m256i*>(tail)), bad, lowest, packed);
        memcpy(out, packed, (size_t)std::min(32u, n_codes - i) / 4 * 3);
    }
    if (lowest_out) {
        __m128i m = _mm_min_epu8(_mm256_castsi256_si128(lowest), _mm256_extracti128_si256(lowest, 1));
        m = _mm_min_epu8(m, _mm_srli_si128(m, 8));
        m = _mm_min_epu8(m, _mm_srli_si128(m, 4));
        m = _mm_min_epu8(m, _mm_srli_si128(m, 2));
        m = _mm_min_epu8(m, _mm_srli_si128(m, 1));
        *lowest_out = (uint8_t)_mm_extract_epi8(m, 0);
    }
    return _mm256_testz_si256(bad, bad) != 0;
}
#endif
#if defined(__x86_64__)
// 64 characters -> 48 bytes per step: masked load ('!' where the line has ended), two multiply-adds that merge four codes
// into 24 bits, one byte permute across the vector that drops every fourth byte, masked store of exactly the bytes due
BC_AVX512 inline bool pack_qual6_avx512(const uint8_t* c, uint32_t n_valid, uint32_t n_codes, uint8_t* out, uint8_t* lowest_out) {
    const __m512i bang = _mm512_set1_epi8('!'), k33 = _mm512_set1_epi8(33), k63 = _mm512_set1_epi8(63), k62 = _mm512_set1_epi8(62),
                  kff = _mm512_set1_epi8((char)0xFF), mul1 = _mm512_set1_epi16(0x4001), mul2 = _mm512_set1_epi32(0x10000001);
    alignas(64) static const uint8_t kIdx[64] = {0,  1,  2,  4,  5,  6,  8,  9,  10, 12, 13, 14, 16, 17, 18, 20, 21, 22, 24, 25, 26, 28,
                                                 29, 30, 32, 33, 34, 36, 37, 38, 40, 41, 42, 44, 45, 46, 48, 49, 50, 52, 53, 54, 56, 57,
                                                 58, 60, 61, 62, 0,  0,  0,  0,  0,  0,  0,  0,  0,  0,  0,  0,  0,  0,  0,  0};
    const __m512i idx = _mm512_load_si512(kIdx);
    __m512i lowest = kff;
    uint64_t bad = 0;
    for (uint32_t i = 0; i < n_codes; i += 64, out += 48) {
        const uint32_t have = i < n_valid ? std::min(64u, n_valid - i) : 0u, due = std::min(64u, n_codes - i);
        const __mmask64 k = have >= 64 ? ~0ULL : ((1ULL << have) - 1ULL);
        const __m512i v = _mm512_mask_loadu_epi8(bang, k, c + i);
        const uint64_t mark = _mm512_cmpeq_epi8_mask(v, kff);
        lowest = _mm512_min_epu8(lowest, v);
        __m512i code = _mm512_sub_epi8(v, k33);
        bad |= _mm512_cmpgt_epu8_mask(code, k62) & ~mark;
        code = _mm512_and_si512(_mm512_mask_mov_epi8(code, mark, k63), k63);
        const __m512i m = _mm512_madd_epi16(_mm512_maddubs_epi16(code, mul1), mul2);
        const __m512i packed = _mm512_permutexvar_epi8(idx, m);
        const uint32_t bytes = due / 4 * 3;
        _mm512_mask_storeu_epi8(out, bytes >= 64 ? ~0ULL : ((1ULL << bytes) - 1ULL), packed);
    }
    if (lowest_out) {
        __m256i h = _mm256_min_epu8(_mm512_castsi512_si256(lowest), _mm512_extracti64x4_epi64(lowest, 1));
        __m128i q = _mm_min_epu8(_mm256_castsi256_si128(h), _mm256_extracti128_si256(h, 1));
        q = _mm_min_epu8(q, _mm_srli_si128(q, 8));
        q = _mm_min_epu8(q, _mm_srli_si128(q, 4));
        q = _mm_min_epu8(q, _mm_srli_si128(q, 2));
        q = _mm_min_epu8(q, _mm_srli_si128(q, 1));
        *lowest_out = (uint8_t)_mm_extract_epi8(q, 0);
    }
    return bad == 0;
}
#endif
inline bool pack_qual6(const uint8_t* c, uint32_t n_codes, uint8_t* out) {
#if defined(__x86_64__)
    if (kHaveAvx512) return pack_qual6_avx512(c, n_codes, n_codes, out, nullptr);
    if (kHaveAvx2) return pack_qual6_avx2(c, n_codes, n_codes, out, nullptr);
#endif
    return pack_qual6_scalar(c, n_codes, out);
}

// n_codes characters -> `bits`-bit dictionary codes (bits 4 or 2); false: a character is not in the dictionary
inline bool pack_qual_dict(const uint8_t* c, uint32_t n_codes, const uint8_t* code_of /* 256 entries, 0xFF = absent */, uint32_t bits,
                           uint8_t* out, uint32_t stride) {
    memset(out, 0, stride);
    bool ok = true;
    for (uint32_t i = 0; i < n_codes; i++) {
        uint8_t code = code_of[c[i]];
        if (code == 0xFF) {
            ok = false;
            code = 0;
        }
        out[(i * bits) >> 3] |= (uint8_t)(code << ((i * bits) & 7));
    }
    return ok;
}

// N plane of reads [first, first + count) -> (read, position) pairs
inline void list_ncalls(const uint32_t* nmask, uint32_t W, size_t first, size_t count, std::vector<uint32_t>& reads, std::vector<uint16_t>& pos) {
    for (size_t r = first; r < first + count; r++)
        for (uint32_t w = 0; w < W; w++) {
            uint32_t m = nmask[r * W + w];
            while (m) {
                reads.push_back((uint32_t)r);
                pos.push_back((uint16_t)(32 * w + (uint32_t)__builtin_ctz(m)));
                m &= m - 1;
            }
        }
}

// reads [0, count) of `reads` -> rows [at, at + count) of a wire arena; returns false when a quality character needs
// the 8-bit form (bits == 6 only).  Too long reads as in pack_range (never strict: the ingest chose the geometry).
bool pack_range_wire(uint32_t mrl, const WireLayout& L, unsigned char* arena, const ReadRef* reads, size_t at, size_t count, uint32_t bits) {
    const uint32_t W = L.W, qs = bits ? bc_wire_qual_stride(mrl, bits) : 0;
    uint32_t* lohi = reinterpret_cast<uint32_t*>(arena + L.o_lohi);
    uint16_t* rl = reinterpret_cast<uint16_t*>(arena + L.o_len);
    uint32_t* nm = reinterpret_cast<uint32_t*>(arena + L.o_nm);
    uint8_t* q = arena + L.o_q;
    uint8_t chars[BC_MAX_READ_LEN + 32];
    bool fits = true;
    for (size_t i = 0; i < count; i++) {
        const ReadRef& r = reads[i];
        const size_t row = at + i;
        uint32_t* lo = lohi + row * 2 * W;
        if (r.len > mrl || r.len > 0x7FFF) {
            memset(lo, 0, (size_t)2 * W * 4);
            memset(nm + row * W, 0, (size_t)W * 4);
            rl[row] = (uint16_t)BC_READ_UNSUPPORTED;
            if (bits) {
                memset(chars, '!', L.n_codes);
                if (bits == 6) pack_qual6(chars, L.n_codes, q + row * qs);
                else memcpy(q + row * qs, chars, L.n_codes);
            }
            continue;
        }
        bool other = pack_planes(r.seq, r.len, W, lo, lo + W, nm + row * W);
#if defined(__x86_64__)
        if (bits == 6 && kHaveAvx2 && r.qlen >= r.len) {  // the usual read: codes straight from the text, one pass
            uint8_t lowest = 255;
            const bool ok = kHaveAvx512 ? pack_qual6_avx512(reinterpret_cast<const uint8_t*>(r.qual), r.len, L.n_codes, q + row * qs, &lowest)
                                        : pack_qual6_avx2(reinterpret_cast<const uint8_t*>(r.qual), r.len, L.n_codes, q + row * qs, &lowest);
            if (r.len && lowest < 33) {  // Q13, as pack_one: flagged, never decoded — any representable characters do
                other = true;
                memset(chars, '!', L.n_codes);
                pack_qual6(chars, L.n_codes, q + row * qs);
            } else {
                fits = ok && fits;
            }
            rl[row] = (uint16_t)(r.len | (other ? BC_READ_UNSUPPORTED : 0u));
            continue;
        }
#endif
        if (bits) {
            const uint32_t n = std::min(r.qlen, r.len);
            unsigned char lowest = 255;
            for (uint32_t k = 0; k < n; k++) lowest = std::min<unsigned char>(lowest, (unsigned char)r.qual[k]);
            if (n && lowest < 33) other = true;  // Q13, as pack_one
            qual_chars(r.qual, r.qlen, r.len, L.n_codes, chars);
            if (bits == 6) {
                if (other && lowest < 33) {  // never decoded: any representable characters do
                    memset(chars, '!', L.n_codes);
                }
                fits = pack_qual6(chars, L.n_codes, q + row * qs) && fits;
            } else {
                memcpy(q + row * qs, chars, L.n_codes);
            }
        }
        rl[row] = (uint16_t)(r.len | (other ? BC_READ_UNSUPPORTED : 0u));
    }
    return fits;
}

// A persistent pool of host threads: run(n, f) calls f(0) .. f(n-1) on the workers and on the calling thread and returns
// when all are done.  One pool per run, created at the first ingest and kept (no thread is spawned per block or batch).
class Pool {
  public:
    explicit Pool(unsigned threads) {
        for (unsigned t = 1; t < std::max(1u, threads); t++) workers_.emplace_back([this] { loop(); });
    }
    ~Pool() {
        {
            std::lock_guard<std::mutex> lk(mu_);
            stop_ = true;
        }
        cv_start_.notify_all();
        for (auto& w : workers_) w.join();
    }
    unsigned size() const { return (unsigned)workers_.size() + 1; }
    void run(size_t n, const std::function<void(size_t)>& f) {
        if (n == 0) return;
        if (n == 1 || workers_.empty()) {
            for (size_t i = 0; i < n; i++) f(i);
            return;
        }
        {
            std::lock_guard<std::mutex> lk(mu_);
            fn_ = &f;
            n_ = n;
            next_.store(0);
            busy_ = workers_.size();
            gen_++;
        }
        cv_start_.notify_all();
        work();
        std::unique_lock<std::mutex> lk(mu_);
        cv_done_.wait(lk, [this] { return busy_ == 0; });
        fn_ = nullptr;
    }

  private:
    void work() {
        for (size_t i = next_++; i < n_; i = next_++) (*fn_)(i);
    }
    void loop() {
        uint64_t seen = 0;
        for (;;) {
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_start_.wait(lk, [&] { return stop_ || gen_ != seen; });
                if (stop_) return;
                seen = gen_;
            }
            work();
            std::lock_guard<std::mutex> lk(mu_);
            if (--busy_ == 0) cv_done_.notify_all();
        }
    }
    std::vector<std::thread> workers_;
    std::mutex mu_;
    std::condition_variable cv_start_, cv_done_;
    const std::function<void(size_t)>* fn_ = nullptr;
    size_t n_ = 0, busy_ = 0;
    std::atomic<size_t> next_{0};
    uint64_t gen_ = 0;
    bool stop_ = false;
};

int pack_refs(uint32_t max_read_len, const std::vector<ReadRef>& reads, size_t first, size_t count, uint32_t* planes,
              uint16_t* read_len, uint8_t* qual, unsigned threads) {
    const uint32_t ps = bc_plane_stride(max_read_len), qs = bc_qual_stride(max_read_len);
    std::atomic<int> bad{0};
    auto work = [&](size_t a, size_t b) {
        if (!pack_range(max_read_len, reads.data() + first + a, b - a, planes + a * ps, read_len + a, qual ? qual + a * qs : nullptr, true))
            bad = 1;
    };
    if (threads <= 1 || count < 4096) {
        work(0, count);
    } else {
        std::vector<std::thread> pool;
        const size_t per = (count + threads - 1) / threads;
        for (unsigned t = 0; t < threads; t++) {
            const size_t a = std::min(count, t * per), b = std::min(count, a + per);
            if (a < b) pool.emplace_back(work, a, b);
        }
        for (auto& th : pool) th.join();
    }
    return bad ? BC_EINVAL : BC_OK;
}

// ---- plain FASTQ through mmap: no read() copy at all; host threads split and pack their own slice of the mapping.
// A record can only start at a line that begins with '@' and whose line after next begins with '+': a quality line
// may begin with '@' too, but then the line after next is a sequence line, which never begins with '+'.
const char* find_record_start(const char* p, const char* end) {
    while (p < end) {
        if (*p == '@') {
            const char* l1 = (const char*)memchr(p, '\n', (size_t)(end - p));
            if (!l1) return end;
            const char* l2 = (const char*)memchr(l1 + 1, '\n', (size_t)(end - l1 - 1));
            if (!l2) return end;
            if (l2 + 1 < end && l2[1] == '+') return p;
        }
        const char* nl = (const char*)memchr(p, '\n', (size_t)(end - p));
        if (!nl) return end;
        p = nl + 1;
    }
    return end;
}

// Line ends of [p, end) one after the other.  32 bytes per step: a record's four line ends come out of about ten compares
// instead of four memchr calls (whose set-up dominates on lines this short).
struct NewlineScan {
    const char* next_block;  // first byte not yet compared
    const char* end;
    const char* cur = nullptr;  // block the pending mask belongs to
    uint32_t mask = 0;
    NewlineScan(const char* p, const char* e) : next_block(p), end(e) {}
#if defined(__x86_64__)
    __attribute__((target("avx2"))) const char* next_avx2() {
        while (mask == 0) {
            if (next_block + 32 > end) {  // the last bytes of the range: no load past its end
                if (next_block >= end) return nullptr;
                const char* nl = (const char*)memchr(next_block, '\n', (size_t)(end - next_block));
                next_block = nl ? nl + 1 : end;
                return nl;
            }
            const __m256i v = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(next_block));
            mask = (uint32_t)_mm256_movemask_epi8(_mm256_cmpeq_epi8(v, _mm256_set1_epi8('\n')));
            cur = next_block;
            next_block += 32;
        }
        const char* at = cur + __builtin_ctz(mask);
        mask &= mask - 1;
        return at;
    }
#endif
#if defined(__x86_64__)
    uint64_t mask64 = 0;
    BC_AVX512 const char* next_avx512() {
        while (mask64 == 0) {
            if (next_block >= end) return nullptr;
            const size_t rem = (size_t)(end - next_block);
            const __mmask64 k = rem >= 64 ? ~0ULL : ((1ULL << rem) - 1ULL);  // the masked load never touches a byte beyond the range
            const __m512i v = _mm512_maskz_loadu_epi8(k, next_block);
            mask64 = _mm512_mask_cmpeq_epi8_mask(k, v, _mm512_set1_epi8('\n'));
            cur = next_block;
            next_block += rem >= 64 ? 64 : rem;
        }
        const char* at = cur + __builtin_ctzll(mask64);
        mask64 &= mask64 - 1;
        return at;
    }
#endif
    const char* next() {
#if defined(__x86_64__)
        if (kHaveAvx512) return next_avx512();
        if (kHaveAvx2) return next_avx2();
#endif
        if (next_block >= end) return nullptr;
        const char* nl = (const char*)memchr(next_block, '\n', (size_t)(end - next_block));
        next_block = nl ? nl + 1 : end;
        return nl;
    }
};

#if defined(__x86_64__)
// AVX-512 framing.  Reads of one run mostly share their length, so a sequence line is expected to end where the last one
// did and a quality line where its sequence did: "the byte there is a line end and none comes before it" is three masked
// compares and one well-predicted branch instead of a search whose trip count changes from record to record (the searches
// lost more to mispredicted branches than to the compares).  A line that is not where it was expected is searched for as
// before, so the records framed are exactly those of the generic splitter.  Masked loads never touch a byte beyond `end`.
BC_AVX512 inline const char* find_nl512(const char* p, const char* end) {
    const __m512i nl = _mm512_set1_epi8('\n');
    while (p < end) {
        const size_t rem = (size_t)(end - p);
        const __mmask64 k = rem >= 64 ? ~0ULL : ((1ULL << rem) - 1ULL);
        const uint64_t m = _mm512_mask_cmpeq_epi8_mask(k, _mm512_maskz_loadu_epi8(k, p), nl);
        if (m) return p + __builtin_ctzll(m);
        p += 64;
    }
    return nullptr;
}
BC_AVX512 inline bool has_nl512(const char* p, size_t n) {  // [p, p + n) lies inside the range
    const __m512i nl = _mm512_set1_epi8('\n');
    uint64_t any = 0;
    for (size_t i = 0; i < n; i += 64) {
        const size_t rem = n - i;
        const __mmask64 k = rem >= 64 ? ~0ULL : ((1ULL << rem) - 1ULL);
        any |= _mm512_mask_cmpeq_epi8_mask(k, _mm512_maskz_loadu_epi8(k, p + i), nl);
    }
    return any != 0;
}
BC_AVX512 const char* split_records_avx512(const char* p, const char* end, bool at_eof, std::vector<ReadRef>& out) {
    size_t expect = 0;  // distance from the start of the last sequence line to its line end
    for (;;) {
        if (p >= end) return p;
        const char* nl0 = find_nl512(p, end);
        if (!nl0) return p;
        const char* s = nl0 + 1;
        const char* nl1 = (expect && s + expect < end && s[expect] == '\n' && !has_nl512(s, expect)) ? s + expect : find_nl512(s, end);
        if (!nl1) return p;
        const char* t = nl1 + 1;
        const char* nl2 = find_nl512(t, end);
        if (!nl2) return p;
        const char* u = nl2 + 1;
        const size_t d1 = (size_t)(nl1 - s);
        const char* nl3 = (u + d1 < end && u[d1] == '\n' && !has_nl512(u, d1)) ? u + d1 : (u < end ? find_nl512(u, end) : nullptr);
        const char* stop = nl3;
        if (!nl3) {
            if (at_eof && u < end) stop = end;  // the file's last line may lack its line end
            else return p;
        }
        size_t l1 = d1, l3 = (size_t)(stop - u);
        if (l1 && s[l1 - 1] == '\r') l1--;
        if (l3 && stop[-1] == '\r') l3--;
        out.push_back(ReadRef{s, u, (uint32_t)l1, (uint32_t)l3});
        expect = d1;
        p = nl3 ? nl3 + 1 : stop;
    }
}
#endif

// whole records of [p, end) -> out; returns the position after the last whole record.  at_eof: the last line may lack '\n'.
const char* split_records(const char* p, const char* end, bool at_eof, std::vector<ReadRef>& out) {
#if defined(__x86_64__)
    if (kHaveAvx512) return split_records_avx512(p, end, at_eof, out);
#endif
    NewlineScan scan(p, end);
    for (;;) {
        const char* line[4];
        uint32_t len[4];
        const char* q = p;
        int k = 0;
        for (; k < 4; k++) {
            const char* nl = q < end ? scan.next() : nullptr;
            const char* stop;
            if (nl) stop = nl;
            else if (at_eof && k == 3 && q < end) stop = end;
            else break;
            size_t l = (size_t)(stop - q);
            if (l && stop[-1] == '\r') l--;
            line[k] = q;
            len[k] = (uint32_t)l;
            q = nl ? nl + 1 : stop;
        }
        if (k < 4) return p;
        out.push_back(ReadRef{line[1], line[3], len[1], len[3]});
        p = q;
    }
}

struct MappedFile {
    const char* data = nullptr;
    size_t size = 0;
    int fd = -1;
    ~MappedFile() {
        if (data && size) munmap(const_cast<char*>(data), size);
        if (fd >= 0) close(fd);
    }
    bool open_plain(const char* path, int populate = 0) {  // false: not a plain regular file we can map (gzip, pipe, empty...)
        fd = open(path, O_RDONLY);
        if (fd < 0) return false;
        struct stat st;
        if (fstat(fd, &st) != 0 || !S_ISREG(st.st_mode) || st.st_size < 2) return false;
        void* m = mmap(nullptr, (size_t)st.st_size, PROT_READ, MAP_PRIVATE | (populate == 1 ? MAP_POPULATE : 0), fd, 0);
        if (m == MAP_FAILED) return false;
        data = static_cast<const char*>(m);
        size = (size_t)st.st_size;
        if ((unsigned char)data[0] == 0x1f && (unsigned char)data[1] == 0x8b) return false;  // gzip
        madvise(m, size, MADV_SEQUENTIAL);
        return true;
    }
};

// ---- output ---------------------------------------------------------------------------------------------------

std::string join(const std::vector<std::string>& v) {
    std::string s;
    for (size_t i = 0; i < v.size(); i++) s += (i ? "," : "") + v[i];
    return s;
}

struct DecodedRow {
    std::string sample;           // sample DNA (file sample: reference DNA; raw: captured DNA; none: "barcode")
    std::vector<std::string> dna; // per counted barcode ("" when absent in an enrichment row)
    std::vector<std::string> out; // what is written: ID when a counted file exists, DNA otherwise
    uint64_t count;
};

void decode_rows(const bch_run& run, const bc_ctx* ctx, const bc_table& t, std::vector<DecodedRow>& rows) {
    const uint32_t ns = (uint32_t)run.slots.size();
    const uint32_t stride = BC_MAX_REF_LEN + 1;
    std::vector<int32_t> idx(ns);
    std::vector<char> str((size_t)ns * stride);
    const size_t old = rows.size();
    rows.resize(old + t.n_rows);  // appends: the rows of several contexts (multi-GPU owners) form one table
    for (uint64_t r = 0; r < t.n_rows; r++) {
        const uint32_t mask = t.mask ? t.mask[r] : 0;
        if (bc_key_decode(ctx, t.key_lo[r], t.key_hi ? t.key_hi[r] : 0, mask, 0, idx.data(), str.data(), stride) != BC_OK)
            throw Error("bc_key_decode failed");
        DecodedRow& d = rows[old + r];
        d.count = t.count[r];
        if (run.sample_slot < 0) d.sample = "barcode";
        else if (run.have_sample_file) d.sample = run.slots[run.sample_slot].dna.at((size_t)idx[run.sample_slot]);
        else d.sample = str.data() + (size_t)run.sample_slot * stride;
        d.dna.resize(run.counted.size());
        d.out.resize(run.counted.size());
        for (size_t k = 0; k < run.counted.size(); k++) {
            const int s = run.counted[k];
            if (mask && !(mask & (1u << k))) continue;  // column left empty (info.rs:847-856, 881-893)
            if (run.have_counted_file) {
                d.dna[k] = run.slots[s].dna.at((size_t)idx[s]);
                d.out[k] = run.slots[s].name.at((size_t)idx[s]);
            } else {
                d.dna[k] = d.out[k] = str.data() + (size_t)s * stride;
            }
        }
    }
}

// writes one file; `names` collects "name<TAB>number of barcode rows" (the pair the reference keeps in
// WriteFiles::output_files / output_counts for its stats file, output.rs:143-165)
void write_text(const std::string& dir, const std::string& name, const std::string& text, std::vector<std::string>& names) {
    std::string path = dir;
    if (!path.empty() && path.back() != '/') path.push_back('/');
    path += name;
    std::ofstream out(path, std::ios::binary);
    if (!out) throw Error("cannot create " + path);
    out << text;
    size_t rows = 0;
    for (char c : text) rows += c == '\n';
    names.push_back(name + "\t" + std::to_string(rows ? rows - 1 : 0));
}


// ---- NUMA placement: the pinned staging buffers of a GPU belong on the memory of the socket the GPU hangs off ----------
int numa_node_of_device(int device) {
    char bus[32] = "";
    if (cudaDeviceGetPCIBusId(bus, sizeof bus, device) != cudaSuccess) {
        cudaGetLastError();
        return -1;
    }
    for (char* c = bus; *c; c++) *c = (char)tolower(*c);
    std::ifstream in(std::string("/sys/bus/pci/devices/") + bus + "/numa_node");
    int node = -1;
    if (!(in >> node)) return -1;
    return node;
}

bool cpus_of_node(int node, cpu_set_t* set) {
    std::ifstream in("/sys/devices/system/node/node" + std::to_string(node) + "/cpulist");
    std::string list;
    if (!std::getline(in, list)) return false;
    CPU_ZERO(set);
    size_t i = 0;
    bool any = false;
    while (i < list.size()) {
        size_t j = i;
        while (j < list.size() && list[j] != ',') j++;
        const std::string part = list.substr(i, j - i);
        const size_t dash = part.find('-');
        const int lo = atoi(part.c_str()), hi = dash == std::string::npos ? lo : atoi(part.c_str() + dash + 1);
        for (int c = lo; c <= hi && c < CPU_SETSIZE; c++) {
            CPU_SET(c, set);
            any = true;
        }
        i = j + 1;
    }
    return any;
}

// runs f on a thread confined to the CPUs of `device`'s NUMA node (pages it allocates land on that node); on boxes that
// do not expose the topology f simply runs
void on_device_node(int device, const std::function<void()>& f) {
    const int node = numa_node_of_device(device);
    cpu_set_t set;
    if (node < 0 || !cpus_of_node(node, &set)) {
        f();
        return;
    }
    std::exception_ptr err;
    std::thread t([&] {
        sched_setaffinity(0, sizeof set, &set);
        try {
            f();
        } catch (...) {
            err = std::current_exception();
        }
    });
    t.join();
    if (err) std::rethrow_exception(err);
}

// ---- the records of one block: one vector per splitter slice, seen as one sequence --------------------------------
struct RecSeq {
    const std::vector<std::vector<ReadRef>>* parts = nullptr;
    std::vector<size_t> first;  // first[t] = index of parts[t][0] in the sequence; first[n] = size
    std::vector<uint32_t> longest;  // longest read of each part
    void index(const std::vector<std::vector<ReadRef>>& p, size_t used) {
        parts = &p;
        first.assign(used + 1, 0);
        longest.assign(used, 0);
        for (size_t t = 0; t < used; t++) first[t + 1] = first[t] + p[t].size();
    }
    size_t size() const { return first.back(); }
    size_t part_of(size_t i) const { return (size_t)(std::upper_bound(first.begin(), first.end(), i) - first.begin()) - 1; }
};

// Splits [pos, blk_end) of a mapped plain FASTQ into records on all pool threads: slice t starts at the first record
// start at or after its nominal position and ends where slice t + 1 starts.  The slice that ends at blk_end — whichever
// it is: later slices are empty when the block holds fewer records than threads — may end with an unterminated line
// when the block is the file's last.  Returns the position after the last whole record.
const char* split_block(Pool& pool, const char* pos, const char* blk_end, bool last, std::vector<std::vector<ReadRef>>& parts,
                        RecSeq& seq, size_t min_slice = 65536) {
    const size_t bytes = (size_t)(blk_end - pos);
    const size_t slices = std::max<size_t>(1, std::min<size_t>(pool.size(), bytes / std::max<size_t>(1, min_slice)));
    if (parts.size() < slices) parts.resize(slices);
    std::vector<const char*> start(slices + 1), stop(slices);
    start[0] = pos;
    start[slices] = blk_end;
    const size_t span = bytes / slices;
    pool.run(slices, [&](size_t t) {
        if (t == 0) return;
        const char* nominal = pos + span * t;
        const char* nl = (const char*)memchr(nominal, '\n', (size_t)(blk_end - nominal));
        start[t] = nl ? find_record_start(nl + 1, blk_end) : blk_end;
    });
    for (size_t t = 1; t < slices; t++) start[t] = std::max(start[t], start[t - 1]);
    std::vector<uint32_t> longest(slices, 0);
    pool.run(slices, [&](size_t t) {
        parts[t].clear();
        const char* end = start[t + 1];
        stop[t] = start[t] < end ? split_records(start[t], end, last && end == blk_end, parts[t]) : start[t];
        uint32_t m = 0;
        for (const ReadRef& r : parts[t]) m = std::max(m, r.len);
        longest[t] = m;
    });
    const char* consumed = pos;
    for (size_t t = 0; t < slices; t++) {
        if (start[t] == start[t + 1]) continue;  // empty slice
        if (start[t + 1] != blk_end && stop[t] != start[t + 1])
            throw Error("malformed FASTQ: a record near byte offset " + std::to_string((size_t)(stop[t] - pos)) +
                        " of the current block does not have four lines");
        consumed = stop[t];
    }
    seq.index(parts, slices);
    seq.longest = longest;
    return consumed;
}

// ---- plain FASTQ in ONE pass over the text ------------------------------------------------------------------------------
// split_block + pack read every byte of the file twice (the split leaves the text in DRAM again before the pack comes back
// to it), and on the ingest host both passes run at memory speed.  Here a host thread takes a chunk of the mapping that
// fits its cache (a few hundred records), frames its records, reserves that many rows of the batch with one atomic and packs
// them at once.  A chunk that does not fit the rows left in the batch is packed as far as it goes; its remainder (a text
// range that starts at a record) is the first work of the next batch.  Rows of a batch are therefore in no particular order,
// which the counts do not depend on.
struct TextRange {
    const char* b;
    const char* e;
};
struct PackedRange {  // records of [b, e) went to rows [base, base + n)
    const char* b;
    const char* e;
    size_t base, n;
};
inline const char* after_record(const ReadRef& r, const char* end) {
    const char* p = r.qual + r.qlen;
    if (p < end && *p == '\r') p++;
    if (p < end && *p == '\n') p++;
    return p;
}

class FusedWalk {
  public:
    FusedWalk(Pool& pool, const char* data, size_t size, size_t chunk_bytes)
        : pool_(pool), data_(data), eof_(data + size), chunk_(std::max<size_t>(chunk_bytes, 64)), n_chunks_((size + chunk_ - 1) / chunk_),
          refs_(pool.size()), done_(pool.size()) {}
    bool more() const { return !carried_.empty() || next_chunk_.load() < n_chunks_; }
    const std::vector<std::vector<PackedRange>>& done() const { return done_; }
    // One batch of up to `cap` rows: pack(t, refs, base, n, first_text, end_text) is called on the pool's threads (t = worker slot).
    // Returns the rows filled.
    template <class Pack>
    size_t batch(size_t cap, Pack pack) {
        std::atomic<size_t> cursor{0}, carried_next{0};
        std::vector<TextRange> deferred;
        std::mutex mu;
        std::string error;
        for (auto& d : done_) d.clear();
        pool_.run(pool_.size(), [&](size_t t) {
            std::vector<ReadRef>& R = refs_[t];
            for (;;) {
                if (cursor.load(std::memory_order_relaxed) >= cap) return;
                TextRange it;
                const size_t ci = carried_next.fetch_add(1);
                if (ci < carried_.size()) {
                    it = carried_[ci];
                } else {
                    const size_t c = next_chunk_.fetch_add(1);
                    if (c >= n_chunks_) return;
                    it = TextRange{bound(c), bound(c + 1)};
#ifdef MADV_POPULATE_READ
                    if (populate_ && it.b < it.e) {
                        const uintptr_t a = (uintptr_t)it.b & ~(uintptr_t)4095, z = ((uintptr_t)it.e + 4095) & ~(uintptr_t)4095;
                        madvise((void*)a, z - a, MADV_POPULATE_READ);
                    }
#endif
                }
                if (it.b >= it.e) continue;
                R.clear();
                const char* stop = split_records(it.b, it.e, it.e == eof_, R);
                if (stop != it.e && it.e != eof_) {  // (at the end of the file a trailing partial record is dropped: the reference never posts it)
                    std::lock_guard<std::mutex> lk(mu);
                    error = "malformed FASTQ: a record near byte offset " + std::to_string((size_t)(stop - data_)) + " does not have four lines";
                    cursor.store(cap);
                    return;
                }
                if (R.empty()) continue;
                const size_t n = R.size(), base = cursor.fetch_add(n), take = base >= cap ? 0 : std::min(n, cap - base);
                const char* cut = take == n ? it.e : (take ? after_record(R[take - 1], it.e) : it.b);
                if (take) {
                    pack(t, R.data(), base, take);
                    done_[t].push_back(PackedRange{it.b, cut, base, take});
                }
                if (take < n) {
                    std::lock_guard<std::mutex> lk(mu);
                    deferred.push_back(TextRange{cut, it.e});
                    return;
                }
            }
        });
        if (!error.empty()) throw Error(error);
        const size_t used = std::min(carried_next.load(), carried_.size());
        carried_.erase(carried_.begin(), carried_.begin() + (std::ptrdiff_t)used);
        carried_.insert(carried_.end(), deferred.begin(), deferred.end());
        return std::min(cursor.load(), cap);
    }

  private:
    // chunk c owns the records that START in [bound(c), bound(c + 1)): a boundary is the first record start after the first
    // line end at or after the nominal position, so both neighbours compute the same one
    const char* bound(size_t c) const {
        if (c == 0) return data_;
        if (c >= n_chunks_) return eof_;
        const char* nominal = data_ + c * chunk_;
        const char* nl = (const char*)memchr(nominal, '\n', (size_t)(eof_ - nominal));
        return nl ? find_record_start(nl + 1, eof_) : eof_;
    }
  public:
    bool populate_ = false;

  private:
    Pool& pool_;
    const char* data_;
    const char* eof_;
    size_t chunk_, n_chunks_;
    std::atomic<size_t> next_chunk_{0};
    std::vector<TextRange> carried_;
    std::vector<std::vector<ReadRef>> refs_;
    std::vector<std::vector<PackedRange>> done_;
};

// ---- one ingest: blocks of records -> pinned batches -> bc_submit, round-robin over the contexts ----------------------
struct Ingest {
    bch_run* run;
    bc_ctx* const* ctxs;
    int n_ctx;
    Pool* pool;
    uint32_t mrl0, batch_reads;
    bool with_qual;
    uint64_t total = 0, n_batches = 0;

    bool wire;
    std::vector<std::vector<uint32_t>> call_reads;  // N calls per packing task
    std::vector<std::vector<uint16_t>> call_pos;
    std::vector<size_t> call_first;
    IngestStats st;
    using Clock = std::chrono::steady_clock;
    static double since(Clock::time_point t0) { return std::chrono::duration<double>(Clock::now() - t0).count(); }
    Clock::time_point t_begin;

    Ingest(bch_run* r, bc_ctx* const* c, int n, unsigned threads, uint32_t batch) : run(r), ctxs(c), n_ctx(n), batch_reads(batch) {
        mrl0 = run->cfg.max_read_len;
        with_qual = run->min_quality > 0.0f;
        wire = run->wire_batches;
        t_begin = Clock::now();
        IngestBuffers& I = run->ingest;
        if (!I.pool || I.pool->size() != std::max(1u, threads)) I.pool = std::make_shared<Pool>(threads);
        pool = I.pool.get();
        if (I.batch_reads != batch_reads || I.with_qual != with_qual || I.mrl != mrl0 || (int)I.lanes.size() != n_ctx || I.wire != wire) {
            I.wire = wire;
            I.lanes.clear();
            for (int d = 0; d < n_ctx; d++) I.lanes.emplace_back(new Lane());
            const size_t block_bytes = std::min<size_t>(std::max<size_t>((size_t)batch_reads * (2 * (size_t)mrl0 + 64), 1u << 20), 1u << 30);
            for (FastqBlock& B : I.blocks) B.buf.resize(block_bytes);
            I.batch_reads = batch_reads;
            I.with_qual = with_qual;
            I.mrl = mrl0;
        }
        for (int d = 0; d < n_ctx; d++) {
            Lane& L = *I.lanes[d];
            const int dev = bc_device_of(ctxs[d]);
            if ((wire ? (void*)L.pinned[0].arena : (void*)L.pinned[0].planes) && L.device == dev) continue;
            L.device = dev;
            on_device_node(dev, [&] {  // first touch on the GPU's own socket
                cudaSetDevice(dev);
                for (PinnedBatch& p : L.pinned) {
                    if (wire) p.alloc_wire(WireLayout(batch_reads, mrl0, with_qual).total);
                    else p.alloc(batch_reads, mrl0, with_qual);
                }
            });
            L.cur = L.in_flight = 0;
        }
        for (auto& L : I.lanes) L->in_flight = 0;
    }

    // records [from, seq.size()) of the block, in batches
    void submit(const RecSeq& seq) {
        size_t done = 0;
        const size_t n_rec = seq.size();
        while (done < n_rec) {
            const int d = (int)(n_batches % (uint64_t)n_ctx);
            Lane& L = *run->ingest.lanes[d];
            bc_ctx* ctx = ctxs[d];
            if (L.in_flight >= 2) {  // the buffer we are about to overwrite was handed to the submit before last: only that copy must be over
                const auto t0 = Clock::now();
                if (bc_wait_older_copies(ctx) != BC_OK) throw Error(bc_last_error(ctx));
                st.wait_s += since(t0);
            }
            PinnedBatch& p = L.pinned[L.cur];
            if (wire) {
                const size_t n = submit_wire(seq, done, n_rec, p, ctx);
                total += n;
                done += n;
                n_batches++;
                L.cur ^= 1;
                L.in_flight++;
                if (run->progress) run->progress(total, run->progress_user);
                continue;
            }
            // geometry of this batch: the run's default, or wider when a read of the parts it touches is longer (reads of
            // any length up to BC_MAX_READ_LEN are decoded; the reference has no length limit, input.rs:115-148)
            size_t n = std::min<size_t>(n_rec - done, batch_reads);
            uint32_t longest = 0;
            for (size_t t = seq.part_of(done), e = seq.part_of(done + n - 1); t <= e; t++) longest = std::max(longest, seq.longest[t]);
            uint32_t mrl = mrl0;
            if (longest > mrl0) {
                mrl = std::min<uint32_t>((longest + 31) / 32 * 32, BC_MAX_READ_LEN);
                n = std::min<size_t>(n, p.cap_planes / ((size_t)bc_plane_stride(mrl) * 4));
                if (with_qual) n = std::min<size_t>(n, p.cap_qual / bc_qual_stride(mrl));
                if (n == 0) throw Error("batch buffers too small for a read of " + std::to_string(longest) + " bases: raise --batch-reads");
            }
            const uint32_t ps = bc_plane_stride(mrl), qs = bc_qual_stride(mrl);
            const size_t chunk = 2048, tasks = (n + chunk - 1) / chunk;
            auto t0 = Clock::now();
            pool->run(tasks, [&](size_t k) {
                size_t a = done + k * chunk;
                const size_t b = std::min(done + n, a + chunk);
                while (a < b) {
                    const size_t t = seq.part_of(a);
                    const size_t take = std::min(b, seq.first[t + 1]) - a, dst = a - done;
                    pack_range(mrl, (*seq.parts)[t].data() + (a - seq.first[t]), take, p.planes + dst * ps, p.read_len + dst,
                               with_qual ? p.qual + dst * qs : nullptr, false);
                    a += take;
                }
            });
            bc_batch b{};
            b.n_reads = (uint32_t)n;
            b.plane_stride = ps;
            b.qual_stride = qs;
            b.location = BC_LOC_HOST;
            b.planes = p.planes;
            b.read_len = p.read_len;
            b.qual = with_qual ? p.qual : nullptr;
            st.pack_s += since(t0);
            t0 = Clock::now();
            if (bc_submit(ctx, &b) != BC_OK) throw Error(bc_last_error(ctx));
            st.submit_s += since(t0);
            st.batches++;
            total += n;
            done += n;
            n_batches++;
            L.cur ^= 1;
            L.in_flight++;
            if (run->progress) run->progress(total, run->progress_user);
        }
    }

    // records [done, ...) of the block -> one batch in its transfer form -> bc_submit_wire; returns the reads taken
    size_t submit_wire(const RecSeq& seq, size_t done, size_t n_rec, PinnedBatch& p, bc_ctx* ctx) {
        size_t n = std::min<size_t>(n_rec - done, batch_reads);
        uint32_t longest = 0;
        for (size_t t = seq.part_of(done), e = seq.part_of(done + n - 1); t <= e; t++) longest = std::max(longest, seq.longest[t]);
        uint32_t mrl = mrl0;
        if (longest > mrl0) {  // a wider batch of fewer reads, as on the bc_batch path
            mrl = std::min<uint32_t>((longest + 31) / 32 * 32, BC_MAX_READ_LEN);
            while (n && WireLayout(n, mrl, with_qual).total > p.cap_arena) n = n * 3 / 4;
            if (n == 0) throw Error("batch buffers too small for a read of " + std::to_string(longest) + " bases: raise --batch-reads");
        }
        const WireLayout L(n, mrl, with_qual);
        const size_t chunk = 2048, tasks = (n + chunk - 1) / chunk;
        if (call_reads.size() < tasks) {
            call_reads.resize(tasks);
            call_pos.resize(tasks);
        }
        call_first.assign(tasks + 1, 0);
        const uint32_t* nm = reinterpret_cast<const uint32_t*>(p.arena + L.o_nm);
        std::atomic<int> need8{0};
        auto pack_all = [&](uint32_t bits, bool planes_too) {
            pool->run(tasks, [&](size_t k) {
                size_t a = done + k * chunk;
                const size_t b = std::min(done + n, a + chunk);
                if (planes_too) {
                    call_reads[k].clear();
                    call_pos[k].clear();
                }
                while (a < b) {
                    const size_t t = seq.part_of(a);
                    const size_t take = std::min(b, seq.first[t + 1]) - a, dst = a - done;
                    if (!pack_range_wire(mrl, L, p.arena, (*seq.parts)[t].data() + (a - seq.first[t]), dst, take, bits)) need8 = 1;
                    if (planes_too) list_ncalls(nm, L.W, dst, take, call_reads[k], call_pos[k]);
                    a += take;
                }
            });
        };
        auto t0 = Clock::now();
        uint32_t bits = with_qual ? 6u : 0u;
        pack_all(bits, true);
        if (need8) {  // a quality character beyond '_': this batch goes as plain bytes
            bits = 8;
            pack_all(bits, false);
            st.batches_qual8++;
        }
        for (size_t k = 0; k < tasks; k++) call_first[k + 1] = call_first[k] + call_reads[k].size();
        const size_t n_calls = call_first[tasks];
        const bool list = n_calls <= L.list_cap;
        if (list && n_calls) {
            uint32_t* nr = reinterpret_cast<uint32_t*>(p.arena + L.o_nr);
            uint16_t* np = reinterpret_cast<uint16_t*>(p.arena + L.o_np);
            pool->run(tasks, [&](size_t k) {
                if (call_reads[k].empty()) return;
                memcpy(nr + call_first[k], call_reads[k].data(), call_reads[k].size() * 4);
                memcpy(np + call_first[k], call_pos[k].data(), call_pos[k].size() * 2);
            });
        }
        if (!list) st.batches_dense_n++;
        st.pack_s += since(t0);
        bc_wire_batch wb{};
        wb.n_reads = (uint32_t)n;
        wb.max_read_len = mrl;
        wb.qual_bits = bits;
        wb.qual_stride = bits ? bc_wire_qual_stride(mrl, bits) : 0;
        wb.n_calls = list ? (uint32_t)n_calls : 0;
        wb.lohi = reinterpret_cast<const uint32_t*>(p.arena + L.o_lohi);
        wb.read_len = reinterpret_cast<const uint16_t*>(p.arena + L.o_len);
        wb.nmask = list ? nullptr : nm;
        wb.n_read = reinterpret_cast<const uint32_t*>(p.arena + L.o_nr);
        wb.n_pos = reinterpret_cast<const uint16_t*>(p.arena + L.o_np);
        wb.qual = bits ? p.arena + L.o_q : nullptr;
        t0 = Clock::now();
        if (bc_submit_wire(ctx, &wb) != BC_OK) throw Error(bc_last_error(ctx));
        st.submit_s += since(t0);
        st.batches++;
        return n;
    }

    // a mapped plain FASTQ file in one pass over its text (FusedWalk): rows packed straight into the transfer form.  A batch
    // that meets what the fast form does not hold — a read longer than the run's geometry, a quality character beyond '_' — is
    // framed again from its text ranges and goes through submit(), which packs wider / plainer.
    void run_fused(const char* data, size_t size) {
        const WireLayout L(batch_reads, mrl0, with_qual);
        const size_t est_record = 2 * (size_t)mrl0 + 32;
        const size_t chunk = std::min<size_t>(256u << 10, std::max<size_t>(4096, (size_t)batch_reads * est_record / 64));
        FusedWalk walk(*pool, data, size, chunk);
        walk.populate_ = run->mmap_populate == 2;
        const size_t T = pool->size();
        std::vector<std::vector<uint32_t>> creads(T);
        std::vector<std::vector<uint16_t>> cpos(T);
        std::vector<std::vector<ReadRef>> parts;
        RecSeq seq;
        const uint32_t bits = with_qual ? 6u : 0u;
        while (walk.more()) {
            const int d = (int)(n_batches % (uint64_t)n_ctx);
            Lane& ln = *run->ingest.lanes[d];
            bc_ctx* ctx = ctxs[d];
            if (ln.in_flight >= 2) {  // only the copy out of the buffer about to be rewritten must be over
                const auto t0 = Clock::now();
                if (bc_wait_older_copies(ctx) != BC_OK) throw Error(bc_last_error(ctx));
                st.wait_s += since(t0);
            }
            PinnedBatch& p = ln.pinned[ln.cur];
            const uint32_t* nm = reinterpret_cast<const uint32_t*>(p.arena + L.o_nm);
            std::atomic<int> anomaly{0};
            for (size_t t = 0; t < T; t++) {
                creads[t].clear();
                cpos[t].clear();
            }
            auto t0 = Clock::now();
            const size_t n = walk.batch(batch_reads, [&](size_t t, const ReadRef* r, size_t base, size_t take) {
                for (size_t i = 0; i < take; i++)
                    if (r[i].len > mrl0) anomaly = 1;
                if (!pack_range_wire(mrl0, L, p.arena, r, base, take, bits)) anomaly = 1;
                list_ncalls(nm, L.W, base, take, creads[t], cpos[t]);
            });
            st.pack_s += since(t0);
            if (n == 0) continue;
            if (anomaly) {
                t0 = Clock::now();
                size_t k = 0;
                for (const auto& list : walk.done()) k += list.size();
                if (parts.size() < k) parts.resize(k);
                k = 0;
                for (const auto& list : walk.done())
                    for (const PackedRange& pr : list) {
                        parts[k].clear();
                        split_records(pr.b, pr.e, pr.e == data + size, parts[k]);
                        if (parts[k].size() != pr.n) throw Error("internal: a text range framed differently the second time");
                        k++;
                    }
                seq.index(parts, k);
                for (size_t i = 0; i < k; i++) {
                    uint32_t m = 0;
                    for (const ReadRef& r : parts[i]) m = std::max(m, r.len);
                    seq.longest[i] = m;
                }
                st.split_s += since(t0);
                submit(seq);
                continue;
            }
            t0 = Clock::now();
            size_t n_calls = 0;
            for (size_t t = 0; t < T; t++) n_calls += creads[t].size();
            const bool list = n_calls <= WireLayout(n, mrl0, with_qual).list_cap && n_calls <= L.list_cap;
            if (list && n_calls) {
                uint32_t* nr = reinterpret_cast<uint32_t*>(p.arena + L.o_nr);
                uint16_t* np = reinterpret_cast<uint16_t*>(p.arena + L.o_np);
                size_t at = 0;
                for (size_t t = 0; t < T; t++) {
                    if (creads[t].empty()) continue;
                    memcpy(nr + at, creads[t].data(), creads[t].size() * 4);
                    memcpy(np + at, cpos[t].data(), cpos[t].size() * 2);
                    at += creads[t].size();
                }
            }
            if (!list) st.batches_dense_n++;
            st.pack_s += since(t0);
            bc_wire_batch wb{};
            wb.n_reads = (uint32_t)n;
            wb.max_read_len = mrl0;
            wb.qual_bits = bits;
            wb.qual_stride = bits ? bc_wire_qual_stride(mrl0, bits) : 0;
            wb.n_calls = list ? (uint32_t)n_calls : 0;
            wb.lohi = reinterpret_cast<const uint32_t*>(p.arena + L.o_lohi);
            wb.read_len = reinterpret_cast<const uint16_t*>(p.arena + L.o_len);
            wb.nmask = list ? nullptr : nm;
            wb.n_read = reinterpret_cast<const uint32_t*>(p.arena + L.o_nr);
            wb.n_pos = reinterpret_cast<const uint16_t*>(p.arena + L.o_np);
            wb.qual = bits ? p.arena + L.o_q : nullptr;
            t0 = Clock::now();
            if (bc_submit_wire(ctx, &wb) != BC_OK) throw Error(bc_last_error(ctx));
            st.submit_s += since(t0);
            st.batches++;
            total += n;
            n_batches++;
            ln.cur ^= 1;
            ln.in_flight++;
            if (run->progress) run->progress(total, run->progress_user);
        }
    }

    void finish() {
        const auto t0 = Clock::now();
        for (int d = 0; d < n_ctx; d++)
            if (bc_sync(ctxs[d]) != BC_OK) throw Error(bc_last_error(ctxs[d]));
        st.wait_s += since(t0);
        st.total_s = since(t_begin);
        run->stats = st;
    }
};

void check_fastq_name(const std::string& path) {
    auto ends = [&](const char* suf) {
        const size_t k = strlen(suf);
        return path.size() >= k && path.compare(path.size() - k, k, suf) == 0;
    };
    if (!ends("fastq") && !ends("fastq.gz"))  // input.rs:33-39
        throw Error("This program only works with *.fastq files and *.fastq.gz files.  The latter is still experimental");
}

// input::read_fastq for one or several contexts (throws)
uint64_t ingest_fastq(bch_run* run, bc_ctx* const* ctxs, int n_ctx, const char* fastq_path, unsigned threads, uint32_t batch_reads) {
    if (batch_reads == 0) batch_reads = 1u << 20;
    if (threads == 0) threads = std::max(1u, std::thread::hardware_concurrency());
    check_fastq_name(fastq_path);
    Ingest ing(run, ctxs, n_ctx, threads, batch_reads);
    IngestBuffers& I = run->ingest;
    MappedFile mf;
    const auto t_map = Ingest::Clock::now();
    const bool plain = mf.open_plain(fastq_path, run->mmap_populate);
    ing.st.map_s = Ingest::since(t_map);
    if (plain && run->wire_batches && run->fused_ingest) {
        ing.run_fused(mf.data, mf.size);
        ing.finish();
        return ing.total;
    }
    if (plain) {
        // plain file: no read() copy at all; every host thread splits and packs its own slice of each block of the mapping
        const size_t block_bytes = I.blocks[0].buf.size();
        const char* pos = mf.data;
        const char* const eof = mf.data + mf.size;
        std::vector<std::vector<ReadRef>> parts;
        RecSeq seq;
        while (pos < eof) {
            const char* blk_end = std::min(eof, pos + block_bytes);
            const bool last = blk_end == eof;
            const auto t0 = Ingest::Clock::now();
            const char* consumed = split_block(*ing.pool, pos, blk_end, last, parts, seq);
            ing.st.split_s += Ingest::since(t0);
            if (seq.size() == 0) {
                if (last) break;  // trailing partial record: dropped, as the reference never posts it
                throw Error("FASTQ record longer than the block buffer");
            }
            ing.submit(seq);
            pos = consumed;
        }
        ing.finish();
        return ing.total;
    }
    // gzip / bgzip / anything that cannot be mapped: a reader thread fills and splits block i + 1 while this thread packs block i
    for (FastqBlock& B : I.blocks) B.state = 0;
    FastqStream in(fastq_path, threads);
    std::mutex mu;
    std::condition_variable cv;
    bool abort_reader = false, reader_done = false;
    std::string reader_error;
    std::thread reader([&]() {
        try {
            for (int i = 0;; i ^= 1) {
                FastqBlock& B = I.blocks[i];
                {
                    std::unique_lock<std::mutex> lk(mu);
                    cv.wait(lk, [&] { return B.state == 0 || abort_reader; });
                    if (abort_reader) return;
                }
                const bool more = in.next_block(B, batch_reads);
                std::lock_guard<std::mutex> lk(mu);
                if (!more) {
                    reader_done = true;
                    cv.notify_all();
                    return;
                }
                B.state = 1;
                cv.notify_all();
            }
        } catch (const std::exception& e) {
            std::lock_guard<std::mutex> lk(mu);
            reader_error = e.what();
            reader_done = true;
            cv.notify_all();
        }
    });
    auto stop_reader = [&]() {
        {
            std::lock_guard<std::mutex> lk(mu);
            abort_reader = true;
        }
        cv.notify_all();
        if (reader.joinable()) reader.join();
    };
    try {
        std::vector<std::vector<ReadRef>> parts(1);
        RecSeq seq;
        for (int i = 0;; i ^= 1) {
            FastqBlock& B = I.blocks[i];
            {
                std::unique_lock<std::mutex> lk(mu);
                cv.wait(lk, [&] { return B.state == 1 || reader_done; });
                if (B.state != 1) break;
            }
            parts[0].swap(B.recs);
            seq.index(parts, 1);
            uint32_t m = 0;
            for (const ReadRef& r : parts[0]) m = std::max(m, r.len);
            seq.longest[0] = m;
            ing.submit(seq);
            // the pinned copy of the block's bytes is complete when submit returns (packing is synchronous)
            parts[0].swap(B.recs);
            {
                std::lock_guard<std::mutex> lk(mu);
                B.state = 0;
            }
            cv.notify_all();
        }
    } catch (...) {
        stop_reader();
        throw;
    }
    stop_reader();
    if (!reader_error.empty()) throw Error(reader_error);
    ing.finish();
    return ing.total;
}

}  // namespace

extern "C" {

bch_run* bch_open(const bch_args* args, char* err, int errlen) {
    auto report = [&](const std::string& m) {
        if (err && errlen > 0) snprintf(err, (size_t)errlen, "%s", m.c_str());
    };
    if (!args || !args->format_path) {
        report("format_path is required");
        return nullptr;
    }
    bch_run* run = new bch_run();
    try {
        parse_scheme(*run, slurp(args->format_path, "Failed to open"));
        if (args->sample_barcodes_path) load_samples(*run, args->sample_barcodes_path);
        if (args->counted_barcodes_path) load_counted(*run, args->counted_barcodes_path);
        // caps (info.rs:499-532): flag value, else a fifth of the length, rounded down
        if (run->sample_slot >= 0)
            run->max_sample = args->max_errors_sample >= 0 ? (uint16_t)args->max_errors_sample : (uint16_t)(run->slots[run->sample_slot].len / 5);
        for (uint16_t sz : run->counted_sizes)
            run->max_counted.push_back(args->max_errors_counted_barcode >= 0 ? (uint16_t)args->max_errors_counted_barcode : (uint16_t)(sz / 5));
        run->max_constant = args->max_errors_constant >= 0 ? (uint16_t)args->max_errors_constant : (uint16_t)(run->constant_len / 5);
        run->min_quality = args->min_quality;

        bc_config& c = run->cfg;
        c.abi_version = BC_ABI_VERSION;
        c.template_chars = run->format_string.c_str();
        c.template_len = (uint32_t)run->format_string.size();
        c.region_codes = run->regions_string.c_str();
        c.region_len = (uint32_t)run->regions_string.size();
        if (run->slots.size() > BC_MAX_SLOTS) throw Error("scheme: more than " + std::to_string(BC_MAX_SLOTS) + " barcodes");
        c.n_slots = (uint32_t)run->slots.size();
        size_t counted_k = 0;
        for (size_t s = 0; s < run->slots.size(); s++) {
            SlotInfo& S = run->slots[s];
            for (const std::string& d : S.dna) S.dna_ptrs.push_back(d.c_str());
            bc_slot& o = c.slots[s];
            o.kind = (uint8_t)S.kind;
            o.offset = S.offset;
            o.len = S.len;
            o.max_err = S.kind == 'S' ? run->max_sample : S.kind == 'B' ? run->max_counted[counted_k++] : 0;
            o.n_ref = (uint32_t)S.dna.size();
            o.ref_seqs = S.dna_ptrs.empty() ? nullptr : S.dna_ptrs.data();
        }
        c.max_const_err = run->max_constant;
        c.min_quality = run->min_quality;
        c.max_read_len = args->max_read_len ? std::max<uint32_t>(args->max_read_len, c.template_len)
                                            : std::max<uint32_t>(160, 2 * c.template_len);
        describe(*run);
        return run;
    } catch (const std::exception& e) {
        report(e.what());
        delete run;
        return nullptr;
    }
}

void bch_close(bch_run* run) { delete run; }
int bch_set_option(bch_run* run, const char* name, long long value) {
    if (!run || !name) return BC_EINVAL;
    if (!strcmp(name, "wire_batches")) {
        run->wire_batches = value != 0;
        return BC_OK;
    }
    if (!strcmp(name, "mmap_populate")) {
        run->mmap_populate = (int)value;
        return BC_OK;
    }
    if (!strcmp(name, "fused_ingest")) {
        run->fused_ingest = value != 0;
        return BC_OK;
    }
    if (!strcmp(name, "lean_writer_min_rows") && value >= 0) run->lean_rows = (uint64_t)value;
    else return BC_EINVAL;
    return BC_OK;
}
void bch_set_progress(bch_run* run, bch_progress_fn fn, void* user) {
    if (!run) return;
    run->progress = fn;
    run->progress_user = user;
}
const bc_config* bch_config(const bch_run* run) { return run ? &run->cfg : nullptr; }
const char* bch_describe(const bch_run* run) { return run ? run->description.c_str() : ""; }
uint32_t bch_barcode_num(const bch_run* run) { return run ? (uint32_t)run->counted.size() : 0; }
const char* bch_ref_dna(const bch_run* run, uint32_t slot, uint32_t i) {
    if (!run || slot >= run->slots.size() || i >= run->slots[slot].dna.size()) return nullptr;
    return run->slots[slot].dna[i].c_str();
}
const char* bch_ref_name(const bch_run* run, uint32_t slot, uint32_t i) {
    if (!run || slot >= run->slots.size() || i >= run->slots[slot].name.size()) return nullptr;
    return run->slots[slot].name[i].c_str();
}

int bch_pack(uint32_t max_read_len, uint32_t n, const char* const* seqs, const char* const* quals, uint32_t* planes_out,
             uint16_t* read_len_out, uint8_t* qual_out, unsigned threads) {
    if (!seqs || !planes_out || !read_len_out || (qual_out && !quals)) return BC_EINVAL;
    std::vector<ReadRef> refs(n);
    for (uint32_t i = 0; i < n; i++)
        refs[i] = ReadRef{seqs[i], quals ? quals[i] : nullptr, (uint32_t)strlen(seqs[i]), quals ? (uint32_t)strlen(quals[i]) : 0u};
    return pack_refs(max_read_len, refs, 0, n, planes_out, read_len_out, quals ? qual_out : nullptr, threads);
}

int bch_pack_lines(uint32_t max_read_len, uint32_t n, const char* seq_lines, const char* qual_lines, uint32_t* planes_out,
                   uint16_t* read_len_out, uint8_t* qual_out, unsigned threads) {
    if (!seq_lines || !planes_out || !read_len_out || (qual_out && !qual_lines)) return BC_EINVAL;
    std::vector<ReadRef> refs;
    refs.reserve(n);
    const char* s = seq_lines;
    const char* q = qual_lines;
    for (uint32_t i = 0; i < n; i++) {
        const char* se = strchr(s, '\n');
        const uint32_t sl = se ? (uint32_t)(se - s) : (uint32_t)strlen(s);
        ReadRef r{s, nullptr, sl, 0};
        if (q) {
            const char* qe = strchr(q, '\n');
            r.qual = q;
            r.qlen = qe ? (uint32_t)(qe - q) : (uint32_t)strlen(q);
            q = qe ? qe + 1 : q + r.qlen;
        }
        refs.push_back(r);
        s = se ? se + 1 : s + sl;
    }
    return pack_refs(max_read_len, refs, 0, n, planes_out, read_len_out, qual_lines ? qual_out : nullptr, threads);
}

int bch_set_simd_level(int level) {
    kHaveAvx2 = kCpuAvx2 && level >= 1;
    kHaveAvx512 = kCpuAvx512 && level >= 2;
    return kHaveAvx512 ? 2 : kHaveAvx2 ? 1 : 0;
}

size_t bch_wire_bound(uint32_t n_reads, uint32_t max_read_len, int with_qual) {
    return WireLayout(n_reads, max_read_len, with_qual != 0).total;
}

// a host bc_batch -> its transfer form in `buf`.  qual_bits 0: the narrowest form that holds every quality character.
int bch_wire_from_batch(const bc_batch* in, uint32_t max_read_len, uint32_t qual_bits, void* buf, size_t buf_bytes, bc_wire_batch* out) {
    if (!in || !out || !buf || in->location != BC_LOC_HOST || !in->planes || !in->read_len) return BC_EINVAL;
    const uint32_t W = bc_plane_words(max_read_len);
    if (max_read_len == 0 || max_read_len > BC_MAX_READ_LEN || in->plane_stride != bc_plane_stride(max_read_len)) return BC_EINVAL;
    const bool with_qual = in->qual != nullptr;
    if (with_qual && in->qual_stride != bc_qual_stride(max_read_len)) return BC_EINVAL;
    const size_t n = in->n_reads;
    const WireLayout L(n, max_read_len, with_qual);
    if (buf_bytes < L.total) return BC_ENOMEM;
    unsigned char* arena = static_cast<unsigned char*>(buf);
    uint32_t* lohi = reinterpret_cast<uint32_t*>(arena + L.o_lohi);
    uint16_t* rl = reinterpret_cast<uint16_t*>(arena + L.o_len);
    uint32_t* nm = reinterpret_cast<uint32_t*>(arena + L.o_nm);
    for (size_t r = 0; r < n; r++) {
        const uint32_t* rec = in->planes + r * in->plane_stride;
        memcpy(lohi + r * 2 * W, rec, (size_t)2 * W * 4);
        memcpy(nm + r * W, rec + 2 * W, (size_t)W * 4);
        rl[r] = in->read_len[r];
    }
    std::vector<uint32_t> cr;
    std::vector<uint16_t> cp;
    list_ncalls(nm, W, 0, n, cr, cp);
    const bool list = cr.size() <= L.list_cap;
    if (list && !cr.empty()) {
        memcpy(arena + L.o_nr, cr.data(), cr.size() * 4);
        memcpy(arena + L.o_np, cp.data(), cp.size() * 2);
    }
    memset(out, 0, sizeof *out);
    out->n_reads = in->n_reads;
    out->max_read_len = max_read_len;
    out->n_calls = list ? (uint32_t)cr.size() : 0;
    out->lohi = lohi;
    out->read_len = rl;
    out->nmask = list ? nullptr : nm;
    out->n_read = reinterpret_cast<const uint32_t*>(arena + L.o_nr);
    out->n_pos = reinterpret_cast<const uint16_t*>(arena + L.o_np);
    if (!with_qual) return BC_OK;
    // which characters occur among the first n_codes of every read
    bool seen[256] = {false};
    const uint32_t n_codes = L.n_codes, have = std::min(n_codes, in->qual_stride);
    // (characters at and beyond a read's length are never looked at by the filter: they do not count, and travel as code 0)
    for (size_t r = 0; r < n; r++) {
        const uint8_t* q = in->qual + r * in->qual_stride;
        const uint32_t len = std::min<uint32_t>(in->read_len[r] & 0x7FFFu, have);
        for (uint32_t i = 0; i < len; i++) seen[q[i]] = true;
    }
    uint32_t distinct = 0;
    bool six = true;
    for (int c = 0; c < 256; c++)
        if (seen[c]) {
            distinct++;
            if (c != 0xFF && (c < 33 || c > 95)) six = false;
        }
    uint32_t bits = qual_bits;
    if (bits == 0) bits = distinct <= 4 ? 2 : distinct <= 16 ? 4 : six ? 6 : 8;
    if ((bits == 2 && distinct > 4) || (bits == 4 && distinct > 16) || (bits == 6 && !six) || (bits != 2 && bits != 4 && bits != 6 && bits != 8))
        return BC_EINVAL;
    uint8_t code_of[256];
    memset(code_of, 0xFF, sizeof code_of);
    if (distinct == 0) out->qual_dict[0] = '!';
    if (bits <= 4) {
        uint32_t k = 0;
        for (int c = 0; c < 256; c++)
            if (seen[c]) {
                out->qual_dict[k] = (uint8_t)c;
                code_of[c] = (uint8_t)k++;
            }
    }
    const uint32_t qs = bc_wire_qual_stride(max_read_len, bits);
    uint8_t* qo = arena + L.o_q;
    std::vector<uint8_t> chars(n_codes + 32);
    for (size_t r = 0; r < n; r++) {
        const uint32_t len = std::min<uint32_t>(in->read_len[r] & 0x7FFFu, have);
        memcpy(chars.data(), in->qual + r * in->qual_stride, len);
        memset(chars.data() + len, bits <= 4 ? out->qual_dict[0] : '!', n_codes - len);
        bool ok = true;
        if (bits == 8) memcpy(qo + r * qs, chars.data(), n_codes);
        else if (bits == 6) ok = pack_qual6(chars.data(), n_codes, qo + r * qs);
        else ok = pack_qual_dict(chars.data(), n_codes, code_of, bits, qo + r * qs, qs);
        if (!ok) return BC_EINVAL;
    }
    out->qual_bits = bits;
    out->qual_stride = qs;
    out->qual = qo;
    return BC_OK;
}

int bch_ingest_stats(const bch_run* run, double* seconds, uint64_t* counts) {
    if (!run || !seconds || !counts) return BC_EINVAL;
    const IngestStats& s = run->stats;
    seconds[0] = s.split_s; seconds[1] = s.pack_s; seconds[2] = s.submit_s; seconds[3] = s.wait_s; seconds[4] = s.total_s;
    seconds[5] = s.map_s;
    counts[0] = s.batches; counts[1] = s.batches_qual8; counts[2] = s.batches_dense_n;
    return BC_OK;
}

int bch_scan_fastq(const char* fastq_path, unsigned threads, uint64_t* n_records, uint64_t* n_bases, uint32_t* crc, char* err,
                   int errlen) {
    if (!fastq_path) return BC_EINVAL;
    if (threads == 0) threads = std::max(1u, std::thread::hardware_concurrency());
    try {
        FastqStream in(fastq_path, threads);
        FastqBlock B;
        B.buf.resize(8u << 20);
        uint64_t recs = 0, bases = 0;
        uLong c = crc32(0L, Z_NULL, 0);
        while (in.next_block(B, 1u << 20)) {
            for (const ReadRef& r : B.recs) {
                c = crc32(c, reinterpret_cast<const unsigned char*>(r.seq), r.len);
                c = crc32(c, reinterpret_cast<const unsigned char*>(r.qual), r.qlen);
                bases += r.len;
            }
            recs += B.recs.size();
        }
        if (n_records) *n_records = recs;
        if (n_bases) *n_bases = bases;
        if (crc) *crc = (uint32_t)c;
        return BC_OK;
    } catch (const std::exception& e) {
        if (err && errlen > 0) snprintf(err, (size_t)errlen, "%s", e.what());
        return BC_EINVAL;
    }
}

int bch_split_fastq(const char* fastq_path, unsigned threads, size_t block_bytes, size_t min_slice_bytes, uint64_t* n_records,
                    uint64_t* n_bases, uint32_t* crc, char* err, int errlen) {
    if (!fastq_path) return BC_EINVAL;
    if (threads == 0) threads = std::max(1u, std::thread::hardware_concurrency());
    if (block_bytes == 0) block_bytes = 64u << 20;
    try {
        MappedFile mf;
        if (!mf.open_plain(fastq_path)) throw Error("not a plain, mappable FASTQ file");
        Pool pool(threads);
        std::vector<std::vector<ReadRef>> parts;
        RecSeq seq;
        uint64_t recs = 0, bases = 0;
        uLong c = crc32(0L, Z_NULL, 0);
        const char* pos = mf.data;
        const char* const eof = mf.data + mf.size;
        while (pos < eof) {
            const char* blk_end = std::min(eof, pos + block_bytes);
            const bool last = blk_end == eof;
            const char* consumed = split_block(pool, pos, blk_end, last, parts, seq, min_slice_bytes ? min_slice_bytes : 65536);
            if (seq.size() == 0) {
                if (last) break;
                throw Error("FASTQ record longer than the block buffer");
            }
            for (size_t t = 0; t + 1 < seq.first.size(); t++)
                for (const ReadRef& r : parts[t]) {
                    c = crc32(c, reinterpret_cast<const unsigned char*>(r.seq), r.len);
                    c = crc32(c, reinterpret_cast<const unsigned char*>(r.qual), r.qlen);
                    bases += r.len;
                }
            recs += seq.size();
            pos = consumed;
        }
        if (n_records) *n_records = recs;
        if (n_bases) *n_bases = bases;
        if (crc) *crc = (uint32_t)c;
        return BC_OK;
    } catch (const std::exception& e) {
        if (err && errlen > 0) snprintf(err, (size_t)errlen, "%s", e.what());
        return BC_EINVAL;
    }
}

// host-only test hook of the one-pass walker: batches of `batch_rows` rows, chunks of chunk_bytes; every record must land on
// exactly one row of exactly one batch.  digest = sum over records of crc32(sequence, quality) (the order is not defined).
int bch_walk_fastq(const char* fastq_path, unsigned threads, size_t chunk_bytes, uint32_t batch_rows, uint64_t* n_records, uint64_t* n_bases,
                   uint64_t* digest, uint64_t* n_batches, char* err, int errlen) {
    if (!fastq_path || batch_rows == 0) return BC_EINVAL;
    if (threads == 0) threads = std::max(1u, std::thread::hardware_concurrency());
    try {
        MappedFile mf;
        if (!mf.open_plain(fastq_path)) throw Error("not a plain, mappable FASTQ file");
        Pool pool(threads);
        FusedWalk walk(pool, mf.data, mf.size, chunk_bytes ? chunk_bytes : (256u << 10));
        uint64_t recs = 0, batches = 0;
        std::atomic<uint64_t> bases{0}, dig{0};
        std::vector<uint8_t> hit(batch_rows);
        while (walk.more()) {
            std::fill(hit.begin(), hit.end(), 0);
            const size_t n = walk.batch(batch_rows, [&](size_t, const ReadRef* r, size_t base, size_t take) {
                uint64_t b = 0, d = 0;
                for (size_t i = 0; i < take; i++) {
                    uLong c = crc32(0L, Z_NULL, 0);
                    c = crc32(c, reinterpret_cast<const unsigned char*>(r[i].seq), r[i].len);
                    c = crc32(c, reinterpret_cast<const unsigned char*>(r[i].qual), r[i].qlen);
                    d += (uint64_t)c;
                    b += r[i].len;
                    hit[base + i]++;
                }
                bases += b;
                dig += d;
            });
            for (size_t i = 0; i < batch_rows; i++)
                if (hit[i] != (i < n ? 1 : 0)) throw Error("row " + std::to_string(i) + " of a batch of " + std::to_string(n) + " was packed " + std::to_string((int)hit[i]) + " times");
            size_t ranged = 0;
            for (const auto& list : walk.done())
                for (const PackedRange& pr : list) ranged += pr.n;
            if (ranged != n) throw Error("the packed ranges do not add up to the batch");
            recs += n;
            batches += n ? 1 : 0;
        }
        if (n_records) *n_records = recs;
        if (n_bases) *n_bases = bases.load();
        if (digest) *digest = dig.load();
        if (n_batches) *n_batches = batches;
        return BC_OK;
    } catch (const std::exception& e) {
        if (err && errlen > 0) snprintf(err, (size_t)errlen, "%s", e.what());
        return BC_EINVAL;
    }
}

int bch_count_fastq(bch_run* run, bc_ctx* ctx, const char* fastq_path, unsigned threads, uint32_t batch_reads,
                    uint64_t* total_reads, char* err, int errlen) {
    if (!run || !ctx || !fastq_path) return BC_EINVAL;
    try {
        run->multi_mode = 0;
        const uint64_t total = ingest_fastq(run, &ctx, 1, fastq_path, threads, batch_reads);
        if (total_reads) *total_reads = total;
        return BC_OK;
    } catch (const std::exception& e) {
        if (err && errlen > 0) snprintf(err, (size_t)errlen, "%s", e.what());
        return BC_EINVAL;
    }
}

// Several GPUs in one process (main.rs:69-121 has one worker pool; here one context per GPU): batches go to the contexts in
// turn; after the last one the ranks merge as SURVEY.md section 8(e) prescribes — hashed keys through ONE exchange of the
// records over NVLink peer memory (every context ends up owning a disjoint set of keys), a dense count table by adding
// the tables into the first context.
int bch_count_fastq_multi(bch_run* run, bc_ctx* const* ctxs, int n_ctx, const char* fastq_path, unsigned threads, uint32_t batch_reads,
                          uint64_t* total_reads, char* err, int errlen) {
    if (!run || !ctxs || n_ctx < 1 || n_ctx > 8 || !fastq_path) return BC_EINVAL;
    try {
        run->multi_mode = 0;
        auto ck = [&](int rc, bc_ctx* c) {
            if (rc != BC_OK) throw Error(bc_last_error(c));
        };
        bc_profile prof;
        ck(bc_get_profile(ctxs[0], &prof), ctxs[0]);
        const bool exchange = n_ctx > 1 && prof.deferred_count;
        auto open_all = [&](uint64_t cap) {
            for (int r = 0; r < n_ctx; r++) ck(bc_exchange_disconnect(ctxs[r]), ctxs[r]);
            for (int r = 0; r < n_ctx; r++) ck(bc_exchange_open(ctxs[r], (uint32_t)n_ctx, (uint32_t)r, cap), ctxs[r]);
            for (int r = 0; r < n_ctx; r++) ck(bc_exchange_connect_local(ctxs[r], ctxs), ctxs[r]);
        };
        uint64_t cap = 0;
        if (exchange) {
            // Receive buffers before the first batch, so that every batch's records leave for their owners right after its
            // decode (streamed exchange).  Sized from the file: an owner gets about 1/n of the reads; a guess that turns out
            // too small only costs the bulk exchange below.
            struct stat st;
            uint64_t bytes = stat(fastq_path, &st) == 0 ? (uint64_t)st.st_size : 0;
            const std::string path = fastq_path;
            if (path.size() > 3 && path.compare(path.size() - 3, 3, ".gz") == 0) bytes *= 4;
            const uint64_t est_reads = bytes / (2 * (uint64_t)std::max<uint32_t>(run->cfg.template_len, 20) + 16) + 1024;
            cap = est_reads / (uint64_t)n_ctx * 5 / 4 + 4096;
            // within the exchange's record limit, and at most a quarter of a GPU's free memory for the two receive buffers
            cap = std::min<uint64_t>(cap, 0xFFFFFFE0ULL);
            size_t free_b = 0, total_b = 0;
            if (cudaSetDevice(bc_device_of(ctxs[0])) == cudaSuccess && cudaMemGetInfo(&free_b, &total_b) == cudaSuccess && free_b) {
                const uint64_t per_record = 2 * 8 * (prof.wide_keys ? 2 : 1);
                cap = std::min<uint64_t>(cap, std::max<uint64_t>(free_b / 4 / per_record, 1u << 16));
            } else {
                cudaGetLastError();
            }
            bool reuse = true;
            for (int r = 0; r < n_ctx; r++) reuse = reuse && bc_exchange_capacity(ctxs[r]) >= cap;
            if (reuse) cap = bc_exchange_capacity(ctxs[0]);
            else open_all(cap);
            for (int r = 1; r < n_ctx; r++)
                if (bc_exchange_capacity(ctxs[r]) != cap) {
                    open_all(cap);
                    break;
                }
        }
        const uint64_t total = ingest_fastq(run, ctxs, n_ctx, fastq_path, threads, batch_reads);
        if (total_reads) *total_reads = total;
        if (n_ctx == 1) return BC_OK;
        if (exchange) {
            // n x n matrix of what every context holds (or has streamed) for every owner
            std::vector<std::vector<uint64_t>> sent(n_ctx, std::vector<uint64_t>(n_ctx, 0));
            for (int r = 0; r < n_ctx; r++) ck(bc_exchange_count(ctxs[r], sent[r].data()), ctxs[r]);
            uint64_t need = 1;
            std::vector<uint64_t> received(n_ctx, 0);
            for (int o = 0; o < n_ctx; o++) {
                for (int r = 0; r < n_ctx; r++) received[o] += sent[r][o];
                need = std::max(need, received[o]);
            }
            if (need > cap) {  // an owner's share did not fit: larger buffers, and the whole exchange in bulk from the record buffers
                open_all(need + need / 16);
                for (int r = 0; r < n_ctx; r++) ck(bc_exchange_count(ctxs[r], sent[r].data()), ctxs[r]);
            }
            for (int r = 0; r < n_ctx; r++) {
                std::vector<uint64_t> first(n_ctx, 0);
                for (int o = 0; o < n_ctx; o++)
                    for (int q = 0; q < r; q++) first[o] += sent[q][o];
                ck(bc_exchange_scatter(ctxs[r], first.data()), ctxs[r]);
            }
            for (int r = 0; r < n_ctx; r++) ck(bc_sync(ctxs[r]), ctxs[r]);  // every scatter has landed
            for (int r = 0; r < n_ctx; r++) ck(bc_exchange_finish(ctxs[r], received[r]), ctxs[r]);
            run->multi_mode = 1;
        } else {
            for (int r = 1; r < n_ctx; r++) {
                ck(bc_sync(ctxs[r]), ctxs[r]);
                ck(bc_peer_add(ctxs[0], ctxs[r], BC_ADD_DENSE_COUNTS), ctxs[0]);
            }
            run->multi_mode = 2;
        }
        return BC_OK;
    } catch (const std::exception& e) {
        if (err && errlen > 0) snprintf(err, (size_t)errlen, "%s", e.what());
        return BC_EINVAL;
    }
}

int bch_counters_multi(bc_ctx* const* ctxs, int n_ctx, uint64_t out[BC_N_COUNTERS]) {
    if (!ctxs || n_ctx < 1 || !out) return BC_EINVAL;
    for (int i = 0; i < BC_N_COUNTERS; i++) out[i] = 0;
    for (int r = 0; r < n_ctx; r++) {
        uint64_t c[BC_N_COUNTERS];
        const int rc = bc_get_counters(ctxs[r], c);
        if (rc != BC_OK) return rc;
        for (int i = 0; i < BC_N_COUNTERS; i++) out[i] += c[i];
    }
    return BC_OK;
}

int bch_write_counts(bch_run* run, bc_ctx* ctx, const char* output_dir, const char* prefix, int merge_output, int enrich,
                     char* names_out, int names_len, char* err, int errlen) {
    if (!run || !ctx) return BC_EINVAL;
    const int mode = run->multi_mode;
    run->multi_mode = 0;  // a single context holds everything
    const int rc = bch_write_counts_multi(run, &ctx, 1, output_dir, prefix, merge_output, enrich, names_out, names_len, err, errlen);
    run->multi_mode = mode;
    return rc;
}

int bch_write_counts_multi(bch_run* run, bc_ctx* const* ctxs, int n_ctx, const char* output_dir, const char* prefix, int merge_output,
                           int enrich, char* names_out, int names_len, char* err, int errlen) {
    auto report = [&](const std::string& m) {
        if (err && errlen > 0) snprintf(err, (size_t)errlen, "%s", m.c_str());
    };
    if (!run || !ctxs || n_ctx < 1) return BC_EINVAL;
    bc_ctx* ctx = ctxs[0];
    const int n_src = (n_ctx > 1 && run->multi_mode == 1) ? n_ctx : 1;  // contexts that hold rows of the result
    bc_table full{}, singles{}, doubles{};
    try {
        const std::string dir = output_dir ? output_dir : "./";
        const std::string pre = prefix ? prefix : "";
        const size_t nb = run->counted.size();
        bool merge = merge_output != 0;
        bool do_enrich = enrich != 0 && nb >= 2;  // main.rs:22-25
        // the rows of every context that holds some (pinned memory owned by the contexts, valid until their next bc_finish)
        std::vector<bc_table> tabs((size_t)n_src);
        uint64_t n_total = 0;
        for (int r = 0; r < n_src; r++) {
            if (bc_finish(ctxs[r], &tabs[(size_t)r]) != BC_OK) throw Error(bc_last_error(ctxs[r]));
            n_total += tabs[(size_t)r].n_rows;
        }
        // Large tables are written by write_full_lean (below): text straight from the packed keys on all host threads, no
        // per-row strings or maps.  The merged file needs every sample's count of a compound side by side, i.e. a map
        // over all rows (as in the reference, output.rs:286-300), so --merge-output keeps the map-based path.
        const bool lean = n_total >= run->lean_rows && !(merge && (run->have_sample_file ? run->slots[run->sample_slot].dna.size() > 1 : run->sample_slot >= 0));
        std::vector<DecodedRow> rows;
        if (!lean)
            for (const bc_table& t : tabs) decode_rows(*run, ctx, t, rows);

        // sample list and its order (output.rs:77-97): with a sample file every listed sample gets files (Q16) and the
        // order is by sample ID; otherwise the samples seen, ordered by DNA for reproducibility
        std::vector<std::string> samples, sample_names;
        if (run->have_sample_file) {
            const SlotInfo& S = run->slots[run->sample_slot];
            std::vector<size_t> order(S.dna.size());
            for (size_t i = 0; i < order.size(); i++) order[i] = i;
            std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) { return S.name[a] < S.name[b]; });
            for (size_t i : order) {
                samples.push_back(S.dna[i]);
                sample_names.push_back(S.name[i]);
            }
        } else if (run->sample_slot >= 0) {
            std::set<std::string> seen;
            if (lean) {
                const uint32_t ns = (uint32_t)run->slots.size(), stride = BC_MAX_REF_LEN + 1;
                std::vector<int32_t> idx(ns);
                std::vector<char> str((size_t)ns * stride);
                for (const bc_table& t : tabs)
                    for (uint64_t r = 0; r < t.n_rows; r++) {
                        if (bc_key_decode(ctx, t.key_lo[r], t.key_hi ? t.key_hi[r] : 0, 0, 0, idx.data(), str.data(), stride) != BC_OK)
                            throw Error("bc_key_decode failed");
                        seen.insert(str.data() + (size_t)run->sample_slot * stride);
                    }
            }
            for (const DecodedRow& r : rows) seen.insert(r.sample);
            for (const std::string& s : seen) {
                samples.push_back(s);
                sample_names.push_back(s);
            }
        } else {
            samples.push_back("barcode");
            sample_names.push_back("barcode");
        }
        std::unordered_map<std::string, size_t> sample_pos;
        for (size_t i = 0; i < samples.size(); i++) sample_pos[samples[i]] = i;

        std::string header = nb > 1 ? "Barcode_1" : "Barcode";  // output.rs:184-196
        for (size_t k = 1; k < nb; k++) header += ",Barcode_" + std::to_string(k + 1);
        if (merge && samples.size() == 1) {  // output.rs:105-110
            fprintf(stderr, "Merged file cannot be created without multiple sample barcodes\n");
            merge = false;
        }
        std::string merged_header = header;
        for (const std::string& nme : sample_names) merged_header += "," + nme;
        merged_header += "\n";

        std::vector<std::string> names;
        std::set<std::string> compounds_written;  // one set across Full / Single / Double, as output.rs:39

        // generic emitter for one family of tables (Full, Single or Double): per-sample files + merged file
        auto emit = [&](const std::vector<DecodedRow>& src, bool full_type, const std::string& descriptor) {
            // per sample: written-key -> count ; merged: code -> (written, per-sample counts)
            struct Merged {
                std::string written;
                std::vector<uint64_t> counts;
            };
            std::vector<std::map<std::string, uint64_t>> per_sample_enriched(samples.size());
            std::vector<std::vector<std::string>> per_sample_lines(samples.size());
            std::map<std::string, Merged> merged;
            for (const DecodedRow& r : src) {
                auto sp = sample_pos.find(r.sample);
                if (sp == sample_pos.end()) continue;
                const size_t si = sp->second;
                const std::string written = join(r.out);
                if (full_type) {
                    per_sample_lines[si].push_back(written + "," + std::to_string(r.count));
                    if (merge) {
                        Merged& m = merged[join(r.dna)];  // keyed by the DNA code (output.rs:292)
                        if (m.counts.empty()) {
                            m.written = written;
                            m.counts.assign(samples.size(), 0);
                        }
                        m.counts[si] += r.count;
                    }
                } else {
                    per_sample_enriched[si][written] += r.count;  // keyed by what is written: IDs merge (Q20)
                }
            }
            if (!full_type) {
                for (size_t si = 0; si < samples.size(); si++)
                    for (const auto& kv : per_sample_enriched[si]) {
                        per_sample_lines[si].push_back(kv.first + "," + std::to_string(kv.second));
                        if (merge) {
                            Merged& m = merged[kv.first];
                            if (m.counts.empty()) {
                                m.written = kv.first;
                                m.counts.assign(samples.size(), 0);
                            }
                            m.counts[si] += kv.second;
                        }
                    }
            }
            for (size_t si = 0; si < samples.size(); si++) {
                std::sort(per_sample_lines[si].begin(), per_sample_lines[si].end());
                std::string text = header + ",Count\n";
                for (const std::string& l : per_sample_lines[si]) text += l + "\n";
                write_text(dir, pre + "_" + sample_names[si] + "_counts" + (descriptor.empty() ? "" : "." + descriptor) + ".csv", text, names);
            }
            if (merge) {
                std::vector<std::string> lines;
                for (const auto& kv : merged) {
                    if (!compounds_written.insert(kv.first).second) continue;
                    std::string l = kv.second.written;
                    for (uint64_t c : kv.second.counts) l += "," + std::to_string(c);
                    lines.push_back(l);
                }
                std::sort(lines.begin(), lines.end());
                std::string text = merged_header;
                for (const std::string& l : lines) text += l + "\n";
                write_text(dir, pre + "_counts.all" + (descriptor.empty() ? "" : "." + descriptor) + ".csv", text, names);
            }
        };

        // The Full family of a large table: every host thread turns a chunk of packed rows into text lines, one buffer per
        // sample; a sample's file is then its chunks one after the other.  Files of up to kSortRows rows are still written
        // byte-sorted; larger ones in the order the GPU produced the rows (the reference writes its hash map's order, Q17).
        auto write_full_lean = [&]() {
            const uint32_t ns = (uint32_t)run->slots.size(), stride = BC_MAX_REF_LEN + 1;
            const size_t kChunk = 1u << 18, kSortRows = 4u << 20, n_samples = samples.size();
            std::vector<size_t> pos_of_idx;  // sample reference index -> position in `samples`
            if (run->have_sample_file) {
                const SlotInfo& S = run->slots[run->sample_slot];
                pos_of_idx.resize(S.dna.size());
                for (size_t i = 0; i < S.dna.size(); i++) pos_of_idx[i] = sample_pos.at(S.dna[i]);
            }
            struct Chunk { const bc_table* t; uint64_t a, b; };
            std::vector<Chunk> chunks;
            for (const bc_table& t : tabs)
                for (uint64_t a = 0; a < t.n_rows; a += kChunk) chunks.push_back(Chunk{&t, a, std::min<uint64_t>(t.n_rows, a + kChunk)});
            std::vector<std::vector<std::string>> text(chunks.size(), std::vector<std::string>(n_samples));
            std::vector<std::vector<uint64_t>> lines(chunks.size(), std::vector<uint64_t>(n_samples, 0));
            std::atomic<int> failed{0};
            Pool pool(std::max(1u, std::thread::hardware_concurrency()));
            pool.run(chunks.size(), [&](size_t c) {
                std::vector<int32_t> idx(ns);
                std::vector<char> str((size_t)ns * stride);
                char num[24];
                const bc_table& t = *chunks[c].t;
                for (uint64_t r = chunks[c].a; r < chunks[c].b; r++) {
                    if (bc_key_decode(ctx, t.key_lo[r], t.key_hi ? t.key_hi[r] : 0, 0, 0, idx.data(), str.data(), stride) != BC_OK) {
                        failed = 1;
                        return;
                    }
                    size_t si = 0;
                    if (run->have_sample_file) si = pos_of_idx[(size_t)idx[run->sample_slot]];
                    else if (run->sample_slot >= 0) si = sample_pos.at(str.data() + (size_t)run->sample_slot * stride);
                    std::string& out = text[c][si];
                    for (size_t k = 0; k < nb; k++) {
                        const int s = run->counted[k];
                        if (k) out.push_back(',');
                        if (run->have_counted_file) out += run->slots[s].name[(size_t)idx[s]];
                        else out += str.data() + (size_t)s * stride;
                    }
                    const int len = snprintf(num, sizeof num, ",%llu\n", (unsigned long long)t.count[r]);
                    out.append(num, (size_t)len);
                    lines[c][si]++;
                }
            });
            if (failed) throw Error("bc_key_decode failed");
            for (size_t si = 0; si < n_samples; si++) {
                uint64_t n_lines = 0;
                size_t bytes = 0;
                for (size_t c = 0; c < chunks.size(); c++) {
                    n_lines += lines[c][si];
                    bytes += text[c][si].size();
                }
                const std::string name = pre + "_" + sample_names[si] + "_counts.csv";
                std::string path = dir;
                if (!path.empty() && path.back() != '/') path.push_back('/');
                std::ofstream out(path + name, std::ios::binary);
                if (!out) throw Error("cannot create " + path + name);
                out << header << ",Count\n";
                if (n_lines <= kSortRows) {
                    std::string all;
                    all.reserve(bytes);
                    for (size_t c = 0; c < chunks.size(); c++) {
                        all += text[c][si];
                        std::string().swap(text[c][si]);
                    }
                    std::vector<std::pair<const char*, size_t>> ls;
                    ls.reserve(n_lines);
                    for (size_t p = 0; p < all.size();) {
                        const size_t e = all.find('\n', p);
                        ls.emplace_back(all.data() + p, e - p);
                        p = e + 1;
                    }
                    std::sort(ls.begin(), ls.end(), [](const std::pair<const char*, size_t>& x, const std::pair<const char*, size_t>& y) {
                        const int c = memcmp(x.first, y.first, std::min(x.second, y.second));
                        return c ? c < 0 : x.second < y.second;
                    });
                    for (const auto& l : ls) {
                        out.write(l.first, (std::streamsize)l.second);
                        out.put('\n');
                    }
                } else {
                    for (size_t c = 0; c < chunks.size(); c++) {
                        out.write(text[c][si].data(), (std::streamsize)text[c][si].size());
                        std::string().swap(text[c][si]);
                    }
                }
                if (!out) throw Error("write failed: " + path + name);
                names.push_back(name + "\t" + std::to_string(n_lines));
            }
        };

        if (lean) write_full_lean();
        else emit(rows, true, "");
        if (do_enrich) {
            std::vector<DecodedRow> srows, drows;
            // the owners' marginals add up: on the device when they are dense counter arrays, else row by row in emit()
            uint64_t* dense = nullptr;
            uint64_t n_dense = 0;
            int n_enrich = n_src;
            if (n_src > 1) {
                if (bc_marginals(ctx, &dense, &n_dense) != BC_OK) throw Error(bc_last_error(ctx));
                if (n_dense) {
                    for (int r = 1; r < n_src; r++) {
                        if (bc_marginals(ctxs[r], &dense, &n_dense) != BC_OK) throw Error(bc_last_error(ctxs[r]));
                        if (bc_sync(ctxs[r]) != BC_OK) throw Error(bc_last_error(ctxs[r]));
                        if (bc_peer_add(ctx, ctxs[r], BC_ADD_MARGINALS) != BC_OK) throw Error(bc_last_error(ctx));
                    }
                    n_enrich = 1;
                }
            }
            for (int r = 0; r < n_enrich; r++) {
                if (bc_enrich(ctxs[r], &singles, nb > 2 ? &doubles : nullptr) != BC_OK) throw Error(bc_last_error(ctxs[r]));
                decode_rows(*run, ctx, singles, srows);
                if (nb > 2) decode_rows(*run, ctx, doubles, drows);  // output.rs:176-178
                bc_table_free(&singles);
                bc_table_free(&doubles);
            }
            emit(srows, false, "Single");
            if (nb > 2) emit(drows, false, "Double");
        }
        if (names_out && names_len > 0) {
            std::string joined;
            for (const std::string& n : names) joined += n + "\n";
            snprintf(names_out, (size_t)names_len, "%s", joined.c_str());
        }
        bc_table_free(&full);
        bc_table_free(&singles);
        bc_table_free(&doubles);
        return (int)names.size();
    } catch (const std::exception& e) {
        bc_table_free(&full);
        bc_table_free(&singles);
        bc_table_free(&doubles);
        report(e.what());
        return BC_EINVAL;
    }
}

}  // extern "C"
