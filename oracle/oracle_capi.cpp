// oracle/oracle_capi.cpp — C entry points over the oracle for the pytest/ctypes harness and bench.py's
// cpu_baseline leg.  TEST INFRASTRUCTURE ONLY; never linked into the product library.
#include <chrono>
#include <cstring>
#include <string>

#include "oracle.hpp"

using namespace oracle;

extern "C" {

typedef struct {
    int status;      // oracle::ReadStatus
    long offset;     // scheme start in the original read, -1 if not located
    int repaired;    // 1 if located by the constant-region repair (phase B)
    int has_random;
    char sample[512];
    char barcodes[4096];  // counted barcodes joined with ','
    char random[512];
} orc_outcome;

static void set_err(char* err, int errlen, const std::string& msg) {
    if (err && errlen > 0) {
        strncpy(err, msg.c_str(), static_cast<size_t>(errlen) - 1);
        err[errlen - 1] = 0;
    }
}

// negative max_err_* means "not given" (20 % default).
void* orc_create(const char* format_path, const char* sample_path, const char* counted_path, int max_err_barcode,
                 int max_err_sample, int max_err_constant, float min_quality, int merge, int enrich,
                 const char* outdir, const char* prefix, char* err, int errlen) {
    try {
        Options o;
        if (max_err_barcode >= 0) o.barcodes_errors = static_cast<uint16_t>(max_err_barcode);
        if (max_err_sample >= 0) o.sample_errors = static_cast<uint16_t>(max_err_sample);
        if (max_err_constant >= 0) o.constant_errors = static_cast<uint16_t>(max_err_constant);
        o.min_quality = min_quality;
        o.merge_output = merge != 0;
        o.enrich = enrich != 0;
        if (outdir) o.output_dir = outdir;
        if (prefix) o.prefix = prefix;
        return new Pipeline(format_path, sample_path ? sample_path : "", counted_path ? counted_path : "", o);
    } catch (const std::exception& e) {
        set_err(err, errlen, e.what());
        return nullptr;
    }
}

void orc_destroy(void* h) { delete static_cast<Pipeline*>(h); }

static void fill(const ReadOutcome& r, orc_outcome* out) {
    out->status = r.status;
    out->offset = r.offset;
    out->repaired = r.repaired ? 1 : 0;
    out->has_random = r.has_random ? 1 : 0;
    snprintf(out->sample, sizeof out->sample, "%s", r.sample_barcode.c_str());
    std::string joined;
    for (size_t i = 0; i < r.counted_barcodes.size(); i++) joined += (i ? "," : "") + r.counted_barcodes[i];
    snprintf(out->barcodes, sizeof out->barcodes, "%s", joined.c_str());
    snprintf(out->random, sizeof out->random, "%s", r.random_barcode.c_str());
}

// decode + count (updates the six counters and the results table)
int orc_process_read(void* h, const char* seq, const char* qual, orc_outcome* out) {
    try {
        fill(static_cast<Pipeline*>(h)->process_read(seq, qual), out);
        return 0;
    } catch (const std::exception&) {
        return -1;
    }
}

// decode only (no counters, no table)
int orc_decode_read(void* h, const char* seq, const char* qual, orc_outcome* out) {
    try {
        fill(static_cast<Pipeline*>(h)->decode_read(seq, qual), out);
        return 0;
    } catch (const std::exception&) {
        return -1;
    }
}

// Many reads at once: seqs/quals are '\n'-separated.  statuses may be NULL.  Returns reads processed.
long orc_process_block(void* h, const char* seqs, const char* quals, int* statuses, long max_reads) {
    Pipeline* p = static_cast<Pipeline*>(h);
    long n = 0;
    const char *s = seqs, *q = quals;
    while (*s && n < max_reads) {
        const char* se = strchr(s, '\n');
        const char* qe = strchr(q, '\n');
        std::string seq = se ? std::string(s, se) : std::string(s);
        std::string qual = qe ? std::string(q, qe) : std::string(q);
        ReadOutcome r = p->process_read(seq, qual);
        if (statuses) statuses[n] = r.status;
        n++;
        if (!se) break;
        s = se + 1;
        q = qe ? qe + 1 : q + strlen(q);
    }
    return n;
}

// order: matched, constant_region, sample_barcode, counted_barcode, duplicates, low_quality (info.rs:146-151)
void orc_counters(void* h, unsigned long long out[6]) {
    const SequenceErrors& e = static_cast<Pipeline*>(h)->errors;
    out[0] = e.matched;
    out[1] = e.constant_region;
    out[2] = e.sample_barcode;
    out[3] = e.barcode;
    out[4] = e.duplicates;
    out[5] = e.low_quality;
}

// writes the CSV set; file names joined by '\n' into names (may be NULL).  Returns number of files or -1.
int orc_write_files(void* h, char* names, int names_len, char* err, int errlen) {
    try {
        std::vector<std::string> files = static_cast<Pipeline*>(h)->write_counts_files();
        if (names && names_len > 0) {
            std::string joined;
            for (const auto& f : files) joined += f + "\n";
            snprintf(names, static_cast<size_t>(names_len), "%s", joined.c_str());
        }
        return static_cast<int>(files.size());
    } catch (const std::exception& e) {
        set_err(err, errlen, e.what());
        return -1;
    }
}

// reference-shaped threaded run over a FASTQ file; returns seconds of wall time, <0 on error
double orc_run_fastq(void* h, const char* path, unsigned threads, unsigned long long* total_reads, char* err, int errlen) {
    try {
        auto t0 = std::chrono::steady_clock::now();
        unsigned long long n = static_cast<Pipeline*>(h)->run_fastq(path, threads);
        auto t1 = std::chrono::steady_clock::now();
        if (total_reads) *total_reads = n;
        return std::chrono::duration<double>(t1 - t0).count();
    } catch (const std::exception& e) {
        set_err(err, errlen, e.what());
        return -1.0;
    }
}

// parse.rs:553-593 on an explicit candidate list; returns 1 and writes the winner, or 0
int orc_fix_error(const char* seq, const char* const* candidates, int n, int max_mismatches, char* out, int outlen) {
    std::vector<std::string> c(candidates, candidates + n);
    auto r = fix_error(seq, c, static_cast<uint16_t>(max_mismatches));
    if (!r) return 0;
    snprintf(out, static_cast<size_t>(outlen), "%s", r->c_str());
    return 1;
}

// info.rs:490-543; sample_size<0 = no sample barcode; *_opt<0 = not given.  barcode_out gets n_barcodes caps.
void orc_max_errors(int sample_errors_opt, int sample_size, int barcode_errors_opt, const unsigned short* barcode_sizes,
                    int n_barcodes, int constant_errors_opt, int constant_region_size, int* constant_out,
                    int* sample_out, int* barcode_out) {
    std::optional<uint16_t> se, ss, be, ce;
    if (sample_errors_opt >= 0) se = static_cast<uint16_t>(sample_errors_opt);
    if (sample_size >= 0) ss = static_cast<uint16_t>(sample_size);
    if (barcode_errors_opt >= 0) be = static_cast<uint16_t>(barcode_errors_opt);
    if (constant_errors_opt >= 0) ce = static_cast<uint16_t>(constant_errors_opt);
    MaxSeqErrors m(se, ss, be, std::vector<uint16_t>(barcode_sizes, barcode_sizes + n_barcodes), ce,
                   static_cast<uint16_t>(constant_region_size), 0.f);
    *constant_out = m.constant_region;
    *sample_out = m.sample_barcode;
    for (int i = 0; i < n_barcodes; i++) barcode_out[i] = m.barcode[static_cast<size_t>(i)];
}

// format introspection for tests: format_string, regions_string, caps
int orc_format_info(void* h, char* format_string, char* regions_string, int buflen, int* constant_len, int* barcode_num,
                    int* max_constant, int* max_sample, int* max_barcode /* [barcode_num] */) {
    Pipeline* p = static_cast<Pipeline*>(h);
    snprintf(format_string, static_cast<size_t>(buflen), "%s", p->format.format_string.c_str());
    snprintf(regions_string, static_cast<size_t>(buflen), "%s", p->format.regions_string.c_str());
    *constant_len = p->format.constant_region_length;
    *barcode_num = static_cast<int>(p->format.barcode_num);
    *max_constant = p->max_errors.constant_region;
    *max_sample = p->max_errors.sample_barcode;
    for (size_t i = 0; i < p->max_errors.barcode.size(); i++) max_barcode[i] = p->max_errors.barcode[i];
    return 0;
}

}  // extern "C"
