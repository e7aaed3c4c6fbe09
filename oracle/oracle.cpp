// oracle/oracle.cpp — see oracle.hpp.  TEST INFRASTRUCTURE ONLY (checker + CPU baseline).
// CPU restatement of the reference's per-read decode-and-count path; citations are file:line
// into the reference repository (src/parse.rs, src/info.rs, src/output.rs, src/input.rs, src/main.rs).
#include "oracle.hpp"

#include <zlib.h>

#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstring>
#include <deque>
#include <fstream>
#include <sstream>
#include <stdexcept>
#include <thread>

namespace oracle {

static const size_t npos = std::string::npos;

// ---------------------------------------------------------------- small helpers

static std::string read_whole_file(const std::string& path, const char* what) {
    std::ifstream in(path, std::ios::binary);
    if (!in) throw std::runtime_error(std::string(what) + " " + path);
    std::stringstream ss;
    ss << in.rdbuf();
    return ss.str();
}

// Rust str::lines(): split on '\n', drop one trailing '\r' per line, no final empty line.
static std::vector<std::string> rust_lines(const std::string& text) {
    std::vector<std::string> out;
    size_t pos = 0;
    while (pos < text.size()) {
        size_t nl = text.find('\n', pos);
        std::string line = nl == npos ? text.substr(pos) : text.substr(pos, nl - pos);
        if (!line.empty() && line.back() == '\r') line.pop_back();
        out.push_back(line);
        if (nl == npos) break;
        pos = nl + 1;
    }
    return out;
}

// Rust str::split(','): always yields at least one (possibly empty) field.
static std::vector<std::string> split_commas(const std::string& s) {
    std::vector<std::string> out;
    size_t pos = 0;
    for (;;) {
        size_t c = s.find(',', pos);
        if (c == npos) {
            out.push_back(s.substr(pos));
            break;
        }
        out.push_back(s.substr(pos, c - pos));
        pos = c + 1;
    }
    return out;
}

static std::string thousands(uint64_t v) {  // num_format Locale::en
    std::string d = std::to_string(v), out;
    for (size_t i = 0; i < d.size(); i++) {
        out.push_back(d[i]);
        size_t left = d.size() - 1 - i;
        if (left && left % 3 == 0) out.push_back(',');
    }
    return out;
}

// ---------------------------------------------------------------- SequenceFormat (info.rs:215-310)

static bool is_digit(char c) { return c >= '0' && c <= '9'; }
static bool is_n(char c) { return c == 'N' || c == 'n'; }
static bool is_atgc(char c) {
    switch (c) {
        case 'A': case 'T': case 'G': case 'C': case 'a': case 't': case 'g': case 'c': return true;
        default: return false;
    }
}

SequenceFormat SequenceFormat::parse_format_text(const std::string& file_text) {
    SequenceFormat f;
    // info.rs:218-222: drop '#' lines, concatenate the rest with no separator
    std::string data;
    for (const std::string& line : rust_lines(file_text))
        if (line.empty() || line[0] != '#') data += line;

    // info.rs:232-233: find_iter of (?i)(\{\d+\})|(\[\d+\])|(\(\d+\))|N+|[ATGC]+ ; anything else is skipped
    std::set<std::string> group_names;
    size_t i = 0;
    while (i < data.size()) {
        char c = data[i];
        std::string group_str;
        if (c == '{' || c == '[' || c == '(') {
            char close = c == '{' ? '}' : (c == '[' ? ']' : ')');
            size_t j = i + 1;
            while (j < data.size() && is_digit(data[j])) j++;
            if (j > i + 1 && j < data.size() && data[j] == close) {
                group_str = data.substr(i, j + 1 - i);
                i = j + 1;
            } else {
                i++;
                continue;
            }
        } else if (is_n(c)) {
            size_t j = i;
            while (j < data.size() && is_n(data[j])) j++;
            group_str = data.substr(i, j - i);
            i = j;
        } else if (is_atgc(c)) {
            size_t j = i;
            while (j < data.size() && is_atgc(data[j])) j++;
            group_str = data.substr(i, j - i);
            i = j;
        } else {
            i++;
            continue;
        }

        std::string group_name;  // info.rs:236-249
        if (group_str.find('[') != npos) {
            group_name = "sample";
            f.sample_barcode = true;
        } else if (group_str.find('{') != npos) {
            f.barcode_num += 1;
            group_name = "barcode" + std::to_string(f.barcode_num);
        } else if (group_str.find('(') != npos) {
            group_name = "random";
            f.random_barcode = true;
        }

        if (!group_name.empty()) {  // info.rs:251-286
            unsigned long digits = std::stoul(group_str.substr(1, group_str.size() - 2));
            if (digits > 65535) throw std::runtime_error("format: barcode length does not fit u16");
            FormatPiece p;
            p.kind = FormatPiece::Capture;
            p.name = group_name;
            p.len = digits;
            f.format_regex.push_back(p);
            if (!group_names.insert(group_name).second)  // Regex::new rejects duplicate names (info.rs:308)
                throw std::runtime_error("format: duplicate capture group name " + group_name);
            char push_char = '\0';
            if (group_name == "sample") {
                f.sample_length_option = static_cast<uint16_t>(digits);
                push_char = 'S';
            } else if (group_name.find("barcode") != npos) {
                f.barcode_lengths.push_back(static_cast<uint16_t>(digits));
                push_char = 'B';
            } else if (group_name == "random") {
                push_char = 'R';
            }
            for (unsigned long k = 0; k < digits; k++) {
                f.regions_string.push_back(push_char);
                f.format_string.push_back('N');
            }
        } else if (group_str.find('N') != npos) {  // info.rs:287-295 (upper-case 'N' only)
            FormatPiece p;
            p.kind = FormatPiece::AnyACGT;
            p.len = static_cast<size_t>(std::count(group_str.begin(), group_str.end(), 'N'));
            f.format_regex.push_back(p);
            f.format_string += group_str;  // nothing is pushed to regions_string (Q9)
        } else {  // info.rs:296-305
            FormatPiece p;
            p.kind = FormatPiece::Literal;
            p.literal = group_str;
            for (char& ch : p.literal) ch = static_cast<char>(std::toupper(static_cast<unsigned char>(ch)));
            p.len = p.literal.size();
            f.format_regex.push_back(p);
            f.format_string += group_str;  // case kept (Q11)
            for (size_t k = 0; k < group_str.size(); k++) f.regions_string.push_back('C');
            f.constant_region_length = static_cast<uint16_t>(f.constant_region_length + group_str.size());
        }
    }
    f.length = f.format_string.size();
    return f;
}

SequenceFormat SequenceFormat::parse_format_file(const std::string& path) {
    return parse_format_text(read_whole_file(path, "Failed to open"));
}

// The reference's regex is a fixed-length concatenation of literals, .{n} and [AGCT]{n}; leftmost-first
// matching of such a pattern is "smallest offset at which every piece matches" (parse.rs:92-95,99-102,153-156).
size_t SequenceFormat::regex_find(const std::string& seq) const {
    size_t pat_len = 0;
    for (const FormatPiece& p : format_regex) pat_len += p.len;
    if (seq.size() < pat_len) return npos;
    for (size_t o = 0; o + pat_len <= seq.size(); o++) {
        size_t pos = o;
        bool ok = true;
        for (const FormatPiece& p : format_regex) {
            if (p.kind == FormatPiece::Literal) {
                if (seq.compare(pos, p.len, p.literal) != 0) ok = false;
            } else if (p.kind == FormatPiece::AnyACGT) {
                for (size_t k = 0; k < p.len; k++) {
                    char c = seq[pos + k];
                    if (!(c == 'A' || c == 'G' || c == 'C' || c == 'T')) ok = false;
                }
            } else {
                for (size_t k = 0; k < p.len; k++)
                    if (seq[pos + k] == '\n') ok = false;  // '.' matches anything but newline
            }
            if (!ok) break;
            pos += p.len;
        }
        if (ok) return o;
    }
    return npos;
}

std::string SequenceFormat::display() const {  // info.rs:313-335
    std::string key;
    std::set<char> seen;
    for (char c : regions_string) {
        if (seen.insert(c).second) {
            switch (c) {
                case 'S': key += "\nS: Sample barcode"; break;
                case 'B': key += "\nB: Counted barcode"; break;
                case 'C': key += "\nC: Constant region"; break;
                case 'R': key += "\nR: Random barcode"; break;
                default: break;
            }
        }
    }
    return "-FORMAT-\n" + format_string + "\n" + regions_string + key;
}

// ---------------------------------------------------------------- MaxSeqErrors (info.rs:490-543)

MaxSeqErrors::MaxSeqErrors(std::optional<uint16_t> sample_errors, std::optional<uint16_t> sample_size_opt,
                           std::optional<uint16_t> barcode_errors, std::vector<uint16_t> barcode_sizes_,
                           std::optional<uint16_t> constant_errors, uint16_t constant_region_size_,
                           float min_quality_) {
    if (sample_size_opt) {  // info.rs:503-513
        sample_size = *sample_size_opt;
        sample_barcode = sample_errors ? *sample_errors : static_cast<uint16_t>(*sample_size_opt / 5);
    } else {
        sample_barcode = 0;
    }
    for (uint16_t size : barcode_sizes_)  // info.rs:517-523
        barcode.push_back(barcode_errors ? *barcode_errors : static_cast<uint16_t>(size / 5));
    constant_region = constant_errors ? *constant_errors : static_cast<uint16_t>(constant_region_size_ / 5);
    constant_region_size = constant_region_size_;
    barcode_sizes = std::move(barcode_sizes_);
    min_quality = min_quality_;
}

static std::string debug_vec(const std::vector<uint16_t>& v) {  // Rust {:?} of Vec<u16>
    std::string s = "[";
    for (size_t i = 0; i < v.size(); i++) s += (i ? ", " : "") + std::to_string(v[i]);
    return s + "]";
}

static std::string rust_f32_display(float v) {  // shortest round-trip decimal, as Rust's Display for f32
    char buf[64];
    for (int prec = 1; prec <= 9; prec++) {
        snprintf(buf, sizeof buf, "%.*g", prec, static_cast<double>(v));
        if (std::strtof(buf, nullptr) == v) break;
    }
    return buf;
}

std::string MaxSeqErrors::display() const {
    std::string size_info, err_info;
    if (barcode_sizes.size() > 1) {
        size_info = "Barcode sizes: " + debug_vec(barcode_sizes);
        err_info = "Maximum mismatches allowed per barcode sequence: " + debug_vec(barcode);
    } else {
        size_info = "Barcode size: " + std::to_string(barcode_sizes.empty() ? 0 : barcode_sizes[0]);
        err_info = "Maximum mismatches allowed per barcode sequence: " + std::to_string(barcode.empty() ? 0 : barcode[0]);
    }
    const std::string bar = "--------------------------------------------------------------\n";
    return "-BARCODE INFO-\nConstant region size: " + std::to_string(constant_region_size) +
           "\nMaximum mismatches allowed per sequence: " + std::to_string(constant_region) + "\n" + bar +
           "Sample barcode size: " + std::to_string(sample_size) +
           "\nMaximum mismatches allowed per sequence: " + std::to_string(sample_barcode) + "\n" + bar + size_info +
           "\n" + err_info + "\n" + bar +
           "Minimum allowed average read quality score per barcode: " + rust_f32_display(min_quality) + "\n";
}

// ---------------------------------------------------------------- BarcodeConversions (info.rs:364-456)

void BarcodeConversions::sample_barcode_file_conversion(const std::string& path) {
    std::vector<std::string> lines = rust_lines(read_whole_file(path, "Failed to open"));
    for (size_t i = 1; i < lines.size(); i++) {  // skip(1): header
        std::vector<std::string> fields = split_commas(lines[i]);
        // take(2).collect_tuple(): a pair only if at least two fields exist, else ("","")
        if (fields.size() >= 2)
            samples_barcode_hash[fields[0]] = fields[1];
        else
            samples_barcode_hash[""] = "";
    }
}

void BarcodeConversions::barcode_file_conversion(const std::string& path, size_t barcode_num) {
    std::vector<std::string> lines = rust_lines(read_whole_file(path, "Failed to read"));
    for (size_t k = 0; k < barcode_num; k++) counted_barcodes_hash.emplace_back();
    std::set<size_t> contained;
    for (size_t i = 1; i < lines.size(); i++) {
        std::vector<std::string> fields = split_commas(lines[i]);
        std::string barcode, id, num;
        if (fields.size() >= 3) {
            barcode = fields[0];
            id = fields[1];
            num = fields[2];
        }
        // str::parse::<usize>: optional leading '+', then digits only
        size_t p = (!num.empty() && num[0] == '+') ? 1 : 0;
        bool ok = p < num.size();
        for (size_t q = p; q < num.size(); q++) ok = ok && is_digit(num[q]);
        if (!ok)
            throw std::runtime_error("Third column of barcode file contains something other than an integer: " + num);
        unsigned long long n = std::stoull(num.substr(p));
        if (n == 0) throw std::runtime_error("barcode number 0: attempt to subtract with overflow");
        size_t idx = static_cast<size_t>(n - 1);
        if (idx >= barcode_num) throw std::runtime_error("barcode number beyond the format's counted barcodes");
        contained.insert(idx);
        counted_barcodes_hash[idx][barcode] = id;  // later duplicates overwrite
    }
    std::string missing;
    for (size_t x = 0; x < barcode_num; x++)
        if (!contained.count(x)) missing += (missing.empty() ? "" : ", ") + std::to_string(x);
    if (!missing.empty())
        throw std::runtime_error("Barcode conversion file missing barcode numers [" + missing + "] in the third column");
}

void BarcodeConversions::get_sample_seqs() {
    for (const auto& kv : samples_barcode_hash) sample_seqs.insert(kv.first);
}

void BarcodeConversions::get_barcode_seqs() {
    if (counted_barcodes_hash.empty()) return;
    counted_barcode_seqs.clear();
    for (const auto& hash : counted_barcodes_hash) {
        std::set<std::string> s;
        for (const auto& kv : hash) s.insert(kv.first);
        counted_barcode_seqs.push_back(std::move(s));
    }
}

// ---------------------------------------------------------------- SequenceErrors display (info.rs:141-172)

std::string SequenceErrors::display() const {
    return "Correctly matched sequences: " + thousands(matched) + "\nConstant region mismatches:  " +
           thousands(constant_region) + "\nSample barcode mismatches:   " + thousands(sample_barcode) +
           "\nCounted barcode mismatches:  " + thousands(barcode) + "\nDuplicates:                  " +
           thousands(duplicates) + "\nLow quality barcodes:        " + thousands(low_quality);
}

// ---------------------------------------------------------------- Results (info.rs:678-808)

Results::Results(const std::map<std::string, std::string>& samples_barcode_hash, bool random_barcode,
                 bool sample_barcode) {
    random_mode = random_barcode;
    if (!samples_barcode_hash.empty()) {  // info.rs:698-709: every listed sample gets a (possibly empty) map (Q16)
        for (const auto& kv : samples_barcode_hash) {
            if (random_mode)
                random_hashmap[kv.first];
            else
                count_hashmap[kv.first];
        }
    } else if (!sample_barcode) {  // info.rs:710-719
        if (random_mode)
            random_hashmap["barcode"];
        else
            count_hashmap["barcode"];
    } else {
        sample_conversion_omited = true;  // info.rs:720-724
    }
}

bool Results::add_count(const std::string& sample_barcode, const std::string* random_barcode,
                        const std::string& barcode_string) {
    if (sample_conversion_omited) {  // info.rs:742-757
        if (random_mode)
            random_hashmap[sample_barcode];
        else
            count_hashmap[sample_barcode];
    }
    if (!random_mode) {  // info.rs:761-767
        auto it = count_hashmap.find(sample_barcode);
        if (it != count_hashmap.end()) it->second[barcode_string] += 1;
        // else: the count lands in a temporary clone and is lost (Q15)
        return true;
    }
    const std::string umi = random_barcode ? *random_barcode : std::string();
    auto it = sample_barcode.empty() ? random_hashmap.find("barcode") : random_hashmap.find(sample_barcode);
    if (it != random_hashmap.end()) {  // info.rs:777-791
        auto& barcodes_hashmap = it->second;
        auto e = barcodes_hashmap.find(barcode_string);
        if (e == barcodes_hashmap.end()) {
            barcodes_hashmap[barcode_string].insert(umi);
        } else {
            return e->second.insert(umi).second;
        }
    } else {  // info.rs:792-801
        random_hashmap[sample_barcode][barcode_string].insert(umi);
    }
    return true;
}

// ---------------------------------------------------------------- fix_error (parse.rs:553-593)

// Returns the index of the unique best candidate, or -1.  Iteration order and the early break are kept
// as in the reference so the restatement can be checked line against line.
template <class It>
static long fix_error_index(const std::string& mismatch_seq, It begin, It end, uint16_t mismatches) {
    long best_match = -1;
    unsigned best_mismatch_count = static_cast<unsigned>(mismatches) + 1;
    bool keep = true;
    long idx = 0;
    for (It it = begin; it != end; ++it, ++idx) {
        const std::string& true_seq = *it;
        unsigned mm = 0;
        size_t n = std::min(true_seq.size(), mismatch_seq.size());  // zip stops at the shorter (Q10)
        for (size_t j = 0; j < n; j++) {
            char possible_char = true_seq[j], current_char = mismatch_seq[j];
            if (possible_char != current_char && current_char != 'N' && possible_char != 'N') mm += 1;
            if (mm > best_mismatch_count) break;
        }
        if (mm == best_mismatch_count) keep = false;
        if (mm < best_mismatch_count) {
            keep = true;
            best_mismatch_count = mm;
            best_match = idx;
        }
    }
    return (keep && best_match >= 0) ? best_match : -1;
}

std::optional<std::string> fix_error(const std::string& mismatch_seq, const std::vector<std::string>& possible_seqs,
                                     uint16_t mismatches) {
    long i = fix_error_index(mismatch_seq, possible_seqs.begin(), possible_seqs.end(), mismatches);
    if (i < 0) return std::nullopt;
    return possible_seqs[static_cast<size_t>(i)];
}

std::optional<std::string> fix_error(const std::string& mismatch_seq, const std::set<std::string>& possible_seqs,
                                     uint16_t mismatches) {
    long i = fix_error_index(mismatch_seq, possible_seqs.begin(), possible_seqs.end(), mismatches);
    if (i < 0) return std::nullopt;
    auto it = possible_seqs.begin();
    std::advance(it, i);
    return *it;
}

// ---------------------------------------------------------------- per-read decode (parse.rs:89-163, 270-375, 439-524)

// parse.rs:287-313.  Returns the chosen window index (or -1) and rewrites `sequence`.
static long fix_constant_region(std::string& sequence, const std::string& format_string, uint16_t max_constant_errors) {
    // parse.rs:291: usize subtraction; R < L is undefined in the reference (Q4) — treated as "no windows" here.
    size_t length_diff = sequence.size() >= format_string.size() ? sequence.size() - format_string.size() : 0;
    std::vector<std::string> possible_seqs;
    for (size_t index = 0; index < length_diff; index++)  // exclusive: the last window is never tried (Q3)
        possible_seqs.push_back(sequence.substr(index, format_string.size()));
    long best = fix_error_index(format_string, possible_seqs.begin(), possible_seqs.end(), max_constant_errors);
    if (best >= 0) {  // parse.rs:270-283
        const std::string& best_sequence = possible_seqs[static_cast<size_t>(best)];
        std::string fixed;
        size_t n = std::min(best_sequence.size(), format_string.size());
        for (size_t j = 0; j < n; j++) fixed.push_back(format_string[j] == 'N' ? best_sequence[j] : format_string[j]);
        sequence = fixed;
    } else {
        sequence.clear();
    }
    return best;
}

// parse.rs:323-375
static bool low_quality(const std::string& quality_values, float min_average, const std::string& indicator,
                        size_t start) {
    std::vector<float> scores;
    char previous_type = '\0';
    for (size_t i = 0; start + i < quality_values.size() && i < indicator.size(); i++) {
        // parse.rs:326: `ch as u8 - 33`; release builds wrap below '!' (Q13)
        uint8_t score = static_cast<uint8_t>(static_cast<uint8_t>(quality_values[start + i]) - 33);
        char seq_type = indicator[i];
        if (seq_type != previous_type) {
            if (!scores.empty()) {
                float sum = 0.f;
                for (float s : scores) sum += s;
                float average_score = sum / static_cast<float>(scores.size());
                if (average_score < min_average) return true;
                scores.clear();
            }
            previous_type = seq_type;
            if (seq_type != 'C') scores.assign(1, static_cast<float>(score));
        } else if (seq_type != 'C') {
            scores.push_back(static_cast<float>(score));
        }
    }
    return false;  // the last run is never tested (Q8)
}

ReadOutcome Pipeline::decode_read(std::string sequence, const std::string& quality) const {
    ReadOutcome out;
    // parse.rs:151-163
    if (format.regex_find(sequence) == npos) {
        long best = fix_constant_region(sequence, format.format_string, max_errors.constant_region);
        if (best >= 0) {
            out.repaired = true;
            out.offset = best;
        }
    }
    // parse.rs:92-96
    size_t start = format.regex_find(sequence);
    if (start == npos) {
        out.status = ConstantRegionError;
        out.repaired = false;
        out.offset = -1;
        return out;
    }
    if (!out.repaired) out.offset = static_cast<long>(start);
    // parse.rs:98-119: quality is read at the regex start in the (possibly rewritten) sequence (Q6)
    if (opt.min_quality > 0.0f && low_quality(quality, opt.min_quality, format.regions_string, start)) {
        out.status = LowQuality;
        return out;
    }
    // captures
    std::map<std::string, std::string> caps;
    size_t pos = start;
    for (const FormatPiece& p : format.format_regex) {
        if (p.kind == FormatPiece::Capture) caps[p.name] = sequence.substr(pos, p.len);
        pos += p.len;
    }
    // parse.rs:448-474
    const std::set<std::string>& sample_seqs = conversions.sample_seqs;
    auto s = caps.find("sample");
    if (s != caps.end()) {
        if (sample_seqs.empty() || sample_seqs.count(s->second)) {
            out.sample_barcode = s->second;
        } else {
            auto fixed = fix_error(s->second, sample_seqs, max_errors.sample_barcode);
            if (fixed) {
                out.sample_barcode = *fixed;
            } else {
                out.status = SampleBarcodeError;
                return out;
            }
        }
    } else {
        out.sample_barcode = "barcode";
    }
    // parse.rs:481-507
    const auto& counted = conversions.counted_barcode_seqs;
    for (size_t index = 0; index < format.barcode_num; index++) {
        std::string counted_barcode = caps["barcode" + std::to_string(index + 1)];
        if (!counted.empty() && !counted[index].count(counted_barcode)) {
            auto fixed = fix_error(counted_barcode, counted[index], max_errors.barcode[index]);
            if (fixed) {
                counted_barcode = *fixed;
            } else {
                out.status = CountedBarcodeError;
                out.counted_barcodes.clear();
                return out;
            }
        }
        out.counted_barcodes.push_back(counted_barcode);
    }
    // parse.rs:510-516
    auto r = caps.find("random");
    if (r != caps.end()) {
        out.random_barcode = r->second;
        out.has_random = true;
    }
    out.status = Matched;  // Matched-or-Duplicate is decided by add_count
    return out;
}

static std::string join_commas(const std::vector<std::string>& v) {
    std::string s;
    for (size_t i = 0; i < v.size(); i++) s += (i ? "," : "") + v[i];
    return s;
}

ReadOutcome Pipeline::process_read(const std::string& sequence, const std::string& quality) {
    ReadOutcome out = decode_read(sequence, quality);
    switch (out.status) {  // parse.rs:111,133,138,145 ; parse.rs:60-69
        case ConstantRegionError: errors.constant_region++; break;
        case LowQuality: errors.low_quality++; break;
        case SampleBarcodeError: errors.sample_barcode++; break;
        case CountedBarcodeError: errors.barcode++; break;
        default: {
            bool added = results.add_count(out.sample_barcode, out.has_random ? &out.random_barcode : nullptr,
                                           join_commas(out.counted_barcodes));
            if (added) {
                errors.matched++;
            } else {
                errors.duplicates++;
                out.status = Duplicate;
            }
        }
    }
    return out;
}

// ---------------------------------------------------------------- Pipeline set-up (main.rs:11-65)

Pipeline::Pipeline(const std::string& format_path, const std::string& sample_path, const std::string& counted_path,
                   const Options& o)
    : opt(o) {
    format = SequenceFormat::parse_format_file(format_path);
    if (opt.enrich && format.barcode_num < 2) opt.enrich = false;  // main.rs:22-25
    if (!sample_path.empty()) {                                     // main.rs:30-33
        conversions.sample_barcode_file_conversion(sample_path);
        conversions.get_sample_seqs();
    }
    results = Results(conversions.samples_barcode_hash, format.random_barcode, format.sample_barcode);  // main.rs:36-40
    if (!counted_path.empty()) {                                                                        // main.rs:43-46
        conversions.barcode_file_conversion(counted_path, format.barcode_num);
        conversions.get_barcode_seqs();
    }
    max_errors = MaxSeqErrors(opt.sample_errors, format.sample_length_option, opt.barcodes_errors,
                              format.barcode_lengths, opt.constant_errors, format.constant_region_length,
                              opt.min_quality);  // main.rs:55-63
}

// ---------------------------------------------------------------- threaded run (main.rs:69-121, input.rs:24-148)

namespace {
struct LineSource {  // plain file or gz (zlib's gz* API walks concatenated members like MultiGzDecoder)
    gzFile gz = nullptr;
    bool is_gz = false;
    std::vector<char> buf;
    explicit LineSource(const std::string& path) : buf(1 << 16) {
        auto ends_with = [&](const char* suf) {
            size_t n = strlen(suf);
            return path.size() >= n && path.compare(path.size() - n, n, suf) == 0;
        };
        is_gz = ends_with("fastq.gz");
        if (!is_gz && !ends_with("fastq"))  // input.rs:35-39
            throw std::runtime_error("This program only works with *.fastq files and *.fastq.gz files.");
        gz = gzopen(path.c_str(), "rb");  // transparent for plain files
        if (!gz) throw std::runtime_error("Failed to open file: " + path);
        gzbuffer(gz, 1 << 20);
    }
    ~LineSource() {
        if (gz) gzclose(gz);
    }
    // returns false at EOF; `line` has no trailing '\n'.  The plain path also strips '\r' (BufRead::lines),
    // the gz path keeps it (read_line), as in input.rs:44-47 vs 66-68 (Q21).
    bool next(std::string& line) {
        line.clear();
        for (;;) {
            if (!gzgets(gz, buf.data(), static_cast<int>(buf.size()))) return !line.empty();
            size_t n = strlen(buf.data());
            line.append(buf.data(), n);
            if (n && buf[n - 1] == '\n') {
                line.pop_back();
                if (!is_gz && !line.empty() && line.back() == '\r') line.pop_back();
                return true;
            }
            if (gzeof(gz)) return !line.empty() || n > 0;
        }
    }
};
}  // namespace

uint64_t Pipeline::run_fastq(const std::string& fastq_path, unsigned threads) {
    if (threads < 2) threads = 2;  // threads==1 spawns no worker and hangs in the reference (main.rs:93)
    std::mutex seq_mutex;
    std::deque<std::string> seq;  // Arc<Mutex<VecDeque<String>>> (main.rs:71)
    std::atomic<bool> finished{false};
    std::atomic<uint64_t> c_const{0}, c_sample{0}, c_barcode{0}, c_match{0}, c_dup{0}, c_lowq{0};
    uint64_t total_reads = 0;
    std::string reader_error;

    std::thread reader([&] {  // input.rs:24-89
        try {
            LineSource src(fastq_path);
            std::string line, record;
            int line_num = 0;
            while (src.next(line)) {
                for (;;) {  // input.rs:117-122 back-pressure
                    std::lock_guard<std::mutex> g(seq_mutex);
                    if (seq.size() < 10000) break;
                }
                line_num = line_num == 4 ? 1 : line_num + 1;
                if (line_num == 1) {
                    total_reads++;
                    record = line;
                } else {
                    record.push_back('\n');
                    record += line;
                }
                if (line_num == 4) {
                    std::lock_guard<std::mutex> g(seq_mutex);
                    seq.push_front(record);
                }
            }
        } catch (const std::exception& e) {
            reader_error = e.what();
        }
        finished.store(true);
    });

    std::vector<std::thread> workers;
    for (unsigned t = 1; t < threads; t++) {
        workers.emplace_back([&] {  // parse.rs:53-76
            for (;;) {
                std::string rec;
                bool got = false;
                {
                    std::lock_guard<std::mutex> g(seq_mutex);
                    if (!seq.empty()) {
                        rec = std::move(seq.back());
                        seq.pop_back();
                        got = true;
                    }
                }
                if (!got) {
                    if (finished.load()) {
                        std::lock_guard<std::mutex> g(seq_mutex);  // close the Q22 race: re-check under the lock
                        if (seq.empty()) break;
                    }
                    continue;
                }
                // RawSequenceRead::unpack (parse.rs:260-267)
                size_t a = rec.find('\n');
                size_t b = a == npos ? npos : rec.find('\n', a + 1);
                size_t c = b == npos ? npos : rec.find('\n', b + 1);
                if (c == npos) continue;
                std::string sequence = rec.substr(a + 1, b - a - 1);
                std::string quality = rec.substr(c + 1);
                ReadOutcome out = decode_read(std::move(sequence), quality);
                switch (out.status) {
                    case ConstantRegionError: c_const++; break;
                    case LowQuality: c_lowq++; break;
                    case SampleBarcodeError: c_sample++; break;
                    case CountedBarcodeError: c_barcode++; break;
                    default: {
                        std::string key = join_commas(out.counted_barcodes);
                        bool added;
                        {
                            std::lock_guard<std::mutex> g(results_mutex_);  // the global results mutex (parse.rs:60)
                            added = results.add_count(out.sample_barcode, out.has_random ? &out.random_barcode : nullptr, key);
                        }
                        if (added) c_match++; else c_dup++;
                    }
                }
            }
        });
    }
    reader.join();
    for (auto& w : workers) w.join();
    if (!reader_error.empty()) throw std::runtime_error("Read Fastq error: " + reader_error);
    errors.constant_region += c_const;
    errors.sample_barcode += c_sample;
    errors.barcode += c_barcode;
    errors.matched += c_match;
    errors.duplicates += c_dup;
    errors.low_quality += c_lowq;
    return total_reads;
}

// ---------------------------------------------------------------- output (output.rs:74-485, info.rs:840-904)

namespace {
using CountMap = std::map<std::string, size_t>;

struct Enrichment {  // info.rs:812-904
    std::map<std::string, CountMap> single_hashmap, double_hashmap;
    void add_sample_barcodes(const std::vector<std::string>& samples) {
        for (const auto& s : samples) {
            single_hashmap[s];
            double_hashmap[s];
        }
    }
    void add_single(const std::string& sample_id, const std::string& barcode_string, size_t count) {
        std::vector<std::string> parts = split_commas(barcode_string);
        size_t barcode_num = parts.size();
        for (size_t index = 0; index < barcode_num; index++) {
            std::string key;
            for (size_t x = 0; x < barcode_num; x++) {
                if (x == index) key += parts[index];
                if (x != barcode_num - 1) key.push_back(',');
            }
            auto it = single_hashmap.find(sample_id);
            if (it != single_hashmap.end()) it->second[key] += count;
        }
    }
    void add_double(const std::string& sample_id, const std::string& barcode_string, size_t count) {
        std::vector<std::string> parts = split_commas(barcode_string);
        size_t barcode_num = parts.size();
        for (size_t first = 0; first + 1 < barcode_num; first++) {
            for (size_t add = 1; add < barcode_num - first; add++) {
                std::string key;
                for (size_t col = 0; col < barcode_num; col++) {
                    if (col == first)
                        key += parts[first];
                    else if (col == first + add)
                        key += parts[first + add];
                    if (col != barcode_num - 1) key.push_back(',');
                }
                auto it = double_hashmap.find(sample_id);
                if (it != double_hashmap.end()) it->second[key] += count;
            }
        }
    }
};

enum class EnrichedType { Single, Double, Full };

struct Writer {
    Pipeline& p;
    Enrichment enriched;
    std::set<std::string> compounds_written;  // shared across Full/Single/Double, as in output.rs:39
    std::string merge_text, sample_text;
    std::vector<std::string> output_files;
    bool merge_output;
    explicit Writer(Pipeline& pl) : p(pl), merge_output(pl.opt.merge_output) {}

    std::string sample_name(const std::string& sample_barcode) const {
        if (p.conversions.samples_barcode_hash.empty()) return sample_barcode;
        auto it = p.conversions.samples_barcode_hash.find(sample_barcode);
        return it == p.conversions.samples_barcode_hash.end() ? "barcode" : it->second;
    }
    void sort_samples(std::vector<std::string>& sample_barcodes) const {  // output.rs:91-97
        if (!p.conversions.samples_barcode_hash.empty())
            std::stable_sort(sample_barcodes.begin(), sample_barcodes.end(),
                             [&](const std::string& a, const std::string& b) { return sample_name(a) < sample_name(b); });
    }
    std::string create_header() const {  // output.rs:184-196
        if (p.format.barcode_num > 1) {
            std::string h = "Barcode_1";
            for (size_t n = 1; n < p.format.barcode_num; n++) h += ",Barcode_" + std::to_string(n + 1);
            return h;
        }
        return "Barcode";
    }
    std::string convert_code(const std::string& code) const {  // output.rs:591-599
        std::vector<std::string> parts = split_commas(code);
        std::string out;
        for (size_t i = 0; i < parts.size(); i++)
            out += (i ? "," : "") + p.conversions.counted_barcodes_hash.at(i).at(parts[i]);
        return out;
    }
    size_t full_count(const std::string& sample, const std::string& code, bool* present) const {
        *present = false;
        if (p.results.random_mode) {
            auto s = p.results.random_hashmap.find(sample);
            if (s == p.results.random_hashmap.end()) return 0;
            auto e = s->second.find(code);
            if (e == s->second.end()) return 0;
            *present = true;
            return e->second.size();
        }
        auto s = p.results.count_hashmap.find(sample);
        if (s == p.results.count_hashmap.end()) return 0;
        auto e = s->second.find(code);
        if (e == s->second.end()) return 0;
        *present = true;
        return e->second;
    }
    // output.rs:199-361
    size_t add_counts_string(const std::string& sample_barcode, const std::vector<std::string>& sample_barcodes,
                             EnrichedType enrichment) {
        std::map<std::string, CountMap> hash_holder;
        std::vector<std::string> codes;
        if (enrichment == EnrichedType::Single) {
            hash_holder = enriched.single_hashmap;
            for (const auto& kv : hash_holder.at(sample_barcode)) codes.push_back(kv.first);
        } else if (enrichment == EnrichedType::Double) {
            hash_holder = enriched.double_hashmap;
            for (const auto& kv : hash_holder.at(sample_barcode)) codes.push_back(kv.first);
        } else if (p.results.random_mode) {
            for (const auto& kv : p.results.random_hashmap.at(sample_barcode)) codes.push_back(kv.first);
        } else {
            for (const auto& kv : p.results.count_hashmap.at(sample_barcode)) codes.push_back(kv.first);
        }
        size_t barcode_num = 0;
        for (const std::string& code : codes) {
            size_t count;
            bool present;
            if (enrichment == EnrichedType::Full)
                count = full_count(sample_barcode, code, &present);
            else
                count = hash_holder.at(sample_barcode).at(code);
            barcode_num++;
            std::string written_barcodes = (enrichment == EnrichedType::Full && !p.conversions.counted_barcodes_hash.empty())
                                               ? convert_code(code)
                                               : code;
            if (merge_output && compounds_written.insert(code).second) {  // output.rs:290-338
                std::string merged_row = written_barcodes;
                for (const std::string& sb : sample_barcodes) {
                    merged_row.push_back(',');
                    size_t c = 0;
                    if (enrichment == EnrichedType::Full) {
                        c = full_count(sb, code, &present);
                    } else {
                        auto& m = hash_holder.at(sb);
                        auto e = m.find(code);
                        c = e == m.end() ? 0 : e->second;
                    }
                    merged_row += std::to_string(c);
                }
                merged_row.push_back('\n');
                merge_text += merged_row;
            }
            sample_text += written_barcodes + "," + std::to_string(count) + "\n";
            if (enrichment == EnrichedType::Full && p.opt.enrich) {  // output.rs:346-353
                enriched.add_single(sample_barcode, written_barcodes, count);
                if (p.format.barcode_num > 2) enriched.add_double(sample_barcode, written_barcodes, count);
            }
        }
        return barcode_num;
    }
    void write_file(const std::string& name, const std::string& text) {
        std::string dir = p.opt.output_dir;
        if (!dir.empty() && dir.back() != '/') dir.push_back('/');
        std::ofstream out(dir + name, std::ios::binary);
        if (!out) throw std::runtime_error("cannot create " + dir + name);
        out << text;
        output_files.push_back(name);
    }
    // Data rows are emitted sorted so runs are reproducible (the reference emits them in ahash order, Q17).
    static std::string sort_rows(const std::string& text) {
        size_t nl = text.find('\n');
        if (nl == npos) return text;
        std::string header = text.substr(0, nl + 1);
        std::vector<std::string> rows;
        size_t pos = nl + 1;
        while (pos < text.size()) {
            size_t e = text.find('\n', pos);
            rows.push_back(text.substr(pos, e - pos));
            pos = e + 1;
        }
        std::sort(rows.begin(), rows.end());
        for (const auto& r : rows) header += r + "\n";
        return header;
    }
    void write_counts_files() {  // output.rs:74-181
        std::vector<std::string> sample_barcodes;
        if (p.results.random_mode)
            for (const auto& kv : p.results.random_hashmap) sample_barcodes.push_back(kv.first);
        else
            for (const auto& kv : p.results.count_hashmap) sample_barcodes.push_back(kv.first);
        if (p.opt.enrich) enriched.add_sample_barcodes(sample_barcodes);
        sort_samples(sample_barcodes);
        std::string header = create_header();
        if (merge_output) {
            if (sample_barcodes.size() == 1) {
                merge_output = false;  // output.rs:105-110 (Q18)
            } else {
                std::string merged_header = header;
                for (const auto& sb : sample_barcodes) merged_header += "," + sample_name(sb);
                merge_text += merged_header + "\n";
            }
        }
        header += ",Count\n";
        for (const auto& sb : sample_barcodes) {
            sample_text += header;
            add_counts_string(sb, sample_barcodes, EnrichedType::Full);
            write_file(p.opt.prefix + "_" + sample_name(sb) + "_counts.csv", sort_rows(sample_text));
            sample_text.clear();
        }
        if (merge_output) {
            write_file(p.opt.prefix + "_counts.all.csv", sort_rows(merge_text));
            merge_text.clear();
        }
        if (p.opt.enrich) {
            write_enriched_files(EnrichedType::Single);
            if (p.format.barcode_num > 2) write_enriched_files(EnrichedType::Double);
        }
    }
    void write_enriched_files(EnrichedType enrichment) {  // output.rs:364-485
        std::vector<std::string> sample_barcodes;
        const auto& src = enrichment == EnrichedType::Single ? enriched.single_hashmap : enriched.double_hashmap;
        for (const auto& kv : src) sample_barcodes.push_back(kv.first);
        sort_samples(sample_barcodes);
        const std::string descriptor = enrichment == EnrichedType::Single ? "Single" : "Double";
        std::string header = create_header();
        if (merge_output) {
            std::string merged_header = header;
            for (const auto& sb : sample_barcodes) merged_header += "," + sample_name(sb);
            merge_text += merged_header + "\n";
        }
        header += ",Count\n";
        for (const auto& sb : sample_barcodes) {
            sample_text += header;
            add_counts_string(sb, sample_barcodes, enrichment);
            write_file(p.opt.prefix + "_" + sample_name(sb) + "_counts." + descriptor + ".csv", sort_rows(sample_text));
            sample_text.clear();
        }
        if (merge_output) {
            write_file(p.opt.prefix + "_counts.all." + descriptor + ".csv", sort_rows(merge_text));
            merge_text.clear();
        }
    }
};
}  // namespace

std::vector<std::string> Pipeline::write_counts_files() {
    Writer w(*this);
    w.write_counts_files();
    return w.output_files;
}

}  // namespace oracle
