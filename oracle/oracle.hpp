// oracle/oracle.hpp — CPU restatement of NGS-Barcode-Count's per-read decode-and-count path.
//
// TEST INFRASTRUCTURE ONLY.  Nothing in the product path (ngs-barcode-count_b200/) may include,
// link or execute this.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs use it, and there only as the checker / CPU baseline.
//
// Pinning status: the reference is Rust and no Rust toolchain exists in this image, so the
// reference itself cannot be run.  The oracle is pinned against the only known-answer tests the
// reference holds for this path (the four rustdoc KATs: parse.rs:540-551, info.rs:559-563,
// 583-587, 607-611) and against an independent, line-by-line Python mirror of the Rust code
// (tests/golden/make_golden.py, which drives Python's `re` with the regex string that
// info.rs:263-298 builds).  Everything beyond the four KATs is therefore "parity unpinned by the
// reference's own tests" (see DESIGN.md §3).
//
// The code is deliberately string based and shaped like the reference (windows as strings,
// char-by-char fix_error, nested maps) so it can be reviewed against parse.rs / info.rs / output.rs.
// Citations are file:line into the reference repository.
#pragma once
#include <cstdint>
#include <map>
#include <mutex>
#include <optional>
#include <set>
#include <string>
#include <vector>

namespace oracle {

// One token of the format file == one piece of the reference's regex (info.rs:233-306).
struct FormatPiece {
    enum Kind { Capture, AnyACGT, Literal } kind;
    std::string name;     // capture group name: "sample", "barcode<k>", "random"
    std::string literal;  // upper-cased constant text (info.rs:298)
    size_t len = 0;
};

// info.rs:176-187
struct SequenceFormat {
    std::string format_string;   // 'N' at barcode and format-N positions, constants verbatim
    std::string regions_string;  // S/B/R per barcode base, C per constant base, nothing for format-N (Q9)
    size_t length = 0;
    uint16_t constant_region_length = 0;
    std::vector<FormatPiece> format_regex;  // fixed-length pattern standing in for regex::Regex
    size_t barcode_num = 0;
    std::vector<uint16_t> barcode_lengths;
    std::optional<uint16_t> sample_length_option;
    bool random_barcode = false;
    bool sample_barcode = false;

    static SequenceFormat parse_format_text(const std::string& file_text);  // info.rs:215-310
    static SequenceFormat parse_format_file(const std::string& path);
    // leftmost match offset of the fixed-length pattern, or npos (regex find/is_match/captures)
    size_t regex_find(const std::string& seq) const;
    std::string display() const;  // info.rs:313-335
};

// info.rs:461-616
struct MaxSeqErrors {
    uint16_t constant_region = 0, constant_region_size = 0;
    uint16_t sample_barcode = 0, sample_size = 0;
    std::vector<uint16_t> barcode, barcode_sizes;
    float min_quality = 0.f;
    MaxSeqErrors() = default;
    MaxSeqErrors(std::optional<uint16_t> sample_errors, std::optional<uint16_t> sample_size_opt,
                 std::optional<uint16_t> barcode_errors, std::vector<uint16_t> barcode_sizes_,
                 std::optional<uint16_t> constant_errors, uint16_t constant_region_size_, float min_quality_);
    std::string display() const;  // info.rs:618-659
};

// info.rs:338-457.  std::map stands in for ahash maps: only iteration order differs (Q17).
struct BarcodeConversions {
    std::map<std::string, std::string> samples_barcode_hash;
    std::set<std::string> sample_seqs;
    std::vector<std::map<std::string, std::string>> counted_barcodes_hash;
    std::vector<std::set<std::string>> counted_barcode_seqs;
    void sample_barcode_file_conversion(const std::string& path);            // info.rs:364-381
    void barcode_file_conversion(const std::string& path, size_t barcode_num);  // info.rs:390-433
    void get_sample_seqs();                                                  // info.rs:435-441
    void get_barcode_seqs();                                                 // info.rs:444-456
};

// info.rs:16-23 — the six QC counters (AtomicU32 in the reference: wraps at 2^32 there).
struct SequenceErrors {
    uint64_t constant_region = 0, sample_barcode = 0, barcode = 0, matched = 0, duplicates = 0, low_quality = 0;
    std::string display() const;  // info.rs:141-172
};

// info.rs:661-809
struct Results {
    bool random_mode = false;
    std::map<std::string, std::map<std::string, std::set<std::string>>> random_hashmap;
    std::map<std::string, std::map<std::string, size_t>> count_hashmap;
    bool sample_conversion_omited = false;
    Results() = default;
    Results(const std::map<std::string, std::string>& samples_barcode_hash, bool random_barcode, bool sample_barcode);
    bool add_count(const std::string& sample_barcode, const std::string* random_barcode,
                   const std::string& barcode_string);  // info.rs:735-808
};

// parse.rs:553-593
std::optional<std::string> fix_error(const std::string& mismatch_seq, const std::vector<std::string>& possible_seqs,
                                     uint16_t mismatches);
std::optional<std::string> fix_error(const std::string& mismatch_seq, const std::set<std::string>& possible_seqs,
                                     uint16_t mismatches);

enum ReadStatus : int {  // exactly one per read (P9)
    Matched = 0,
    Duplicate = 1,
    ConstantRegionError = 2,
    LowQuality = 3,
    SampleBarcodeError = 4,
    CountedBarcodeError = 5,
};

// What happened to one read; everything a parity test wants to look at.
struct ReadOutcome {
    int status = ConstantRegionError;
    long offset = -1;       // located start of the scheme in the ORIGINAL read (-1: not located)
    bool repaired = false;  // located by fix_constant_region (phase B) rather than the regex (phase A)
    std::string sample_barcode;
    std::vector<std::string> counted_barcodes;
    std::string random_barcode;
    bool has_random = false;
};

struct Options {  // arguments.rs:6-20, the subset that shapes the hot path and its outputs
    std::optional<uint16_t> barcodes_errors, sample_errors, constant_errors;
    float min_quality = 0.f;
    bool merge_output = false;
    bool enrich = false;
    std::string output_dir = "./";
    std::string prefix = "oracle";
};

class Pipeline {  // main.rs:11-166 minus clap / timing
  public:
    Pipeline(const std::string& format_path, const std::string& sample_path, const std::string& counted_path,
             const Options& opt);
    // parse.rs:89-148 + parse.rs:55-70 for one read, single threaded
    ReadOutcome process_read(const std::string& sequence, const std::string& quality);
    // decode without touching the shared counters/results (worker side of parse.rs:89-148)
    ReadOutcome decode_read(std::string sequence, const std::string& quality) const;
    // reference-shaped threaded run: 1 reader + (threads-1) workers on a Mutex<VecDeque> (main.rs:69-121)
    // returns number of reads posted
    uint64_t run_fastq(const std::string& fastq_path, unsigned threads);
    // output.rs:74-181 (+ Single/Double, output.rs:364-485); rows written sorted (one of the
    // reference's possible orders, Q17).  Returns the file names written, in reference order.
    std::vector<std::string> write_counts_files();

    SequenceFormat format;
    BarcodeConversions conversions;
    MaxSeqErrors max_errors;
    SequenceErrors errors;
    Results results;
    Options opt;

  private:
    std::mutex results_mutex_;
};

}  // namespace oracle
