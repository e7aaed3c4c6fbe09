// main.cpp — `barcode-count` for B200: the reference's command line (arguments.rs:27-124) over the GPU path.
// Same flags, same input files, same output CSV set; the per-read work runs in the CUDA library (bc_b200.h).
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <string>
#include <thread>
#include <vector>

#include <cuda_runtime_api.h>

#include <zlib.h>

#include "../../../include/bc_host.h"

static void usage() {
    fprintf(stderr,
            "NGS-Barcode-Count (B200 decode-and-count path)\n"
            "Counts barcodes located in sequencing data\n\n"
            "USAGE:\n    barcode-count [FLAGS] [OPTIONS] --fastq <fastq> --sequence-format <format_file>\n\n"
            "FLAGS:\n"
            "    -e, --enrich          Create output files of enrichment for single and double synthons/barcodes\n"
            "    -m, --merge-output    Merge sample output counts into a single file\n"
            "    -h, --help\n    -V, --version\n\n"
            "OPTIONS:\n"
            "    -c, --counted-barcodes <barcode_file>    Counted barcodes file\n"
            "    -o, --output-dir <dir>                   Directory to output the counts to [default: ./]\n"
            "    -f, --fastq <fastq>                      FastQ file\n"
            "    -q, --sequence-format <format_file>      Sequence format file\n"
            "        --max-errors-counted-barcode <n>     Maximimum number of sequence errors allowed within each counted barcode\n"
            "        --max-errors-constant <n>            Maximimum number of sequence errors allowed within constant region\n"
            "        --max-errors-sample <n>              Maximimum number of sequence errors allowed within sample barcode\n"
            "        --min-quality <min>                  Minimum average read quality score per barcode [default: 0]\n"
            "    -p, --prefix <prefix>                    File prefix name [default: today's date]\n"
            "    -s, --sample-barcodes <sample_file>      Sample barcodes file\n"
            "    -t, --threads <threads>                  Number of host threads (FASTQ parse + pack)\n"
            "        --devices <list>                     GPUs to use: 0 | 0,1 | 0-7 | all (reads are sharded over them) [default: 0]\n"
            "        --max-read-length <n>                Longest read in the FASTQ [default: from the first reads of the file]\n"
            "        --batch-reads <n>                    Reads per GPU batch [default: 1048576]\n");
}

static std::string thousands(unsigned long long v) {
    std::string d = std::to_string(v), out;
    for (size_t i = 0; i < d.size(); i++) {
        out.push_back(d[i]);
        const size_t left = d.size() - 1 - i;
        if (left && left % 3 == 0) out.push_back(',');
    }
    return out;
}

static std::string hms(double secs) {
    const long ms = (long)(secs * 1000.0 + 0.5);
    char buf[96];
    snprintf(buf, sizeof buf, "%ld hours, %ld minutes, %ld.%03ld seconds", ms / 3600000, (ms / 60000) % 60, (ms / 1000) % 60, ms % 1000);
    return buf;
}

// Longest sequence line among the first records of the FASTQ (plain or gzip; up to 8 MB of text are looked at), so that
// --max-read-length rarely has to be given by hand.  0 when the file cannot be read.
static unsigned probe_read_len(const std::string& path) {
    gzFile gz = gzopen(path.c_str(), "rb");
    if (!gz) return 0;
    std::string buf(8u << 20, '\0');
    const int got = gzread(gz, &buf[0], (unsigned)buf.size());
    gzclose(gz);
    if (got <= 0) return 0;
    unsigned longest = 0, line_no = 0;
    size_t start = 0;
    for (size_t i = 0; i < (size_t)got; i++) {
        if (buf[i] != '\n') continue;
        if (line_no % 4 == 1) {
            size_t len = i - start;
            if (len && buf[i - 1] == '\r') len--;
            if (len > longest) longest = (unsigned)len;
        }
        line_no++;
        start = i + 1;
    }
    return longest;
}

// The reference's command line is clap's (arguments.rs:27-124): long options take "--name value" or "--name=value", short
// ones "-t 8", "-t8" or "-t=8", and short flags bundle ("-me", "-met8").  argv -> one token per option / value.
static std::vector<std::string> clap_tokens(int argc, char** argv) {
    static const std::string short_with_value = "fqsctop";
    std::vector<std::string> out;
    bool only_values = false;
    for (int i = 1; i < argc; i++) {
        const std::string a = argv[i];
        if (only_values || a.size() < 2 || a[0] != '-') {
            out.push_back(a);
        } else if (a == "--") {
            only_values = true;
        } else if (a[1] == '-') {
            const size_t eq = a.find('=');
            if (eq == std::string::npos) {
                out.push_back(a);
            } else {
                out.push_back(a.substr(0, eq));
                out.push_back(a.substr(eq + 1));
            }
        } else {
            for (size_t k = 1; k < a.size(); k++) {
                out.push_back(std::string("-") + a[k]);
                if (short_with_value.find(a[k]) != std::string::npos) {
                    if (k + 1 < a.size()) out.push_back(a.substr(k + 1 + (a[k + 1] == '=' ? 1 : 0)));
                    break;
                }
            }
        }
    }
    return out;
}

// "0", "0,2,3", "0-3", "all"
static bool parse_devices(const std::string& spec, std::vector<int>& out) {
    out.clear();
    if (spec == "all") {
        int n = 0;
        if (cudaGetDeviceCount(&n) != cudaSuccess || n < 1) return false;
        for (int d = 0; d < n; d++) out.push_back(d);
        return true;
    }
    size_t i = 0;
    while (i < spec.size()) {
        size_t j = spec.find(',', i);
        if (j == std::string::npos) j = spec.size();
        const std::string part = spec.substr(i, j - i);
        const size_t dash = part.find('-');
        char* end = nullptr;
        const long lo = strtol(part.c_str(), &end, 10);
        long hi = lo;
        if (dash != std::string::npos) hi = strtol(part.c_str() + dash + 1, &end, 10);
        if (part.empty() || lo < 0 || hi < lo || hi > 63) return false;
        for (long d = lo; d <= hi; d++) out.push_back((int)d);
        i = j + 1;
    }
    return !out.empty() && out.size() <= 8;
}

int main(int argc, char** argv) {
    const auto t0 = std::chrono::steady_clock::now();
    std::string fastq, format, samples, counted, outdir = "./", prefix;
    int max_b = -1, max_s = -1, max_c = -1;
    std::vector<int> devices{0};
    float min_quality = 0.f;
    bool merge = false, enrich = false;
    unsigned threads = std::thread::hardware_concurrency();
    unsigned max_read_len = 0, batch_reads = 1u << 20;
    {
        char buf[32];
        const time_t now = time(nullptr);
        strftime(buf, sizeof buf, "%Y-%m-%d", localtime(&now));
        prefix = buf;
    }
    const std::vector<std::string> tok = clap_tokens(argc, argv);
    for (size_t i = 0; i < tok.size(); i++) {
        const std::string& a = tok[i];
        auto val = [&]() -> const char* {
            if (i + 1 >= tok.size()) {
                fprintf(stderr, "error: The argument '%s' requires a value but none was supplied\n", a.c_str());
                exit(2);
            }
            return tok[++i].c_str();
        };
        auto num = [&](const char* what) -> int {
            const char* v = val();
            char* end = nullptr;
            const long x = strtol(v, &end, 10);
            if (!*v || *end || x < 0 || x > 65535) {
                fprintf(stderr, "Error: Unable to convert %s to an integer\n", what);
                exit(1);
            }
            return (int)x;
        };
        if (a == "-f" || a == "--fastq") fastq = val();
        else if (a == "-q" || a == "--sequence-format") format = val();
        else if (a == "-s" || a == "--sample-barcodes") samples = val();
        else if (a == "-c" || a == "--counted-barcodes") counted = val();
        else if (a == "-t" || a == "--threads") threads = (unsigned)num("threads");
        else if (a == "-o" || a == "--output-dir") outdir = val();
        else if (a == "-p" || a == "--prefix") prefix = val();
        else if (a == "-m" || a == "--merge-output") merge = true;
        else if (a == "-e" || a == "--enrich") enrich = true;
        else if (a == "--max-errors-counted-barcode") max_b = num("maximum barcode errors");
        else if (a == "--max-errors-sample") max_s = num("maximum sample errors");
        else if (a == "--max-errors-constant") max_c = num("maximum constant errors");
        else if (a == "--min-quality") {
            const char* v = val();
            char* end = nullptr;
            min_quality = strtof(v, &end);
            if (!*v || *end) {
                fprintf(stderr, "Error: Unable to convert min score to a float\n");
                return 1;
            }
        } else if (a == "--device" || a == "--devices") {
            const char* v = val();
            if (!parse_devices(v, devices)) {
                fprintf(stderr, "Error: --devices wants a list like 0,1 or 0-7 (at most 8 GPUs of one box), or 'all'\n");
                return 1;
            }
        } else if (a == "--max-read-length") max_read_len = (unsigned)num("max read length");
        else if (a == "--batch-reads") batch_reads = (unsigned)strtoul(val(), nullptr, 10);
        else if (a == "-h" || a == "--help") { usage(); return 0; }
        else if (a == "-V" || a == "--version") { printf("NGS-Barcode-Count 0.11.1-b200\n"); return 0; }
        else {
            fprintf(stderr, "error: Found argument '%s' which wasn't expected\n", a.c_str());
            usage();
            return 2;
        }
    }
    if (fastq.empty() || format.empty()) {
        fprintf(stderr, "error: The following required arguments were not provided:\n    --fastq <fastq>\n    --sequence-format <format_file>\n");
        return 2;
    }
    char err[2048] = "";
    bch_args args{};
    args.format_path = format.c_str();
    args.sample_barcodes_path = samples.empty() ? nullptr : samples.c_str();
    args.counted_barcodes_path = counted.empty() ? nullptr : counted.c_str();
    args.max_errors_counted_barcode = max_b;
    args.max_errors_sample = max_s;
    args.max_errors_constant = max_c;
    args.min_quality = min_quality;
    if (max_read_len == 0) {  // the default batch geometry: the longest of the first reads, with some head-room; a batch that
        const unsigned seen = probe_read_len(fastq);  // meets a longer read is packed wider, nothing aborts
        if (seen) max_read_len = std::min(1024u, seen + seen / 8 + 8);
    }
    args.max_read_len = max_read_len;
    bch_run* run = bch_open(&args, err, sizeof err);
    if (!run) {
        fprintf(stderr, "Error: %s\n", err);
        return 1;
    }
    printf("%s\n", bch_describe(run));
    if (enrich && bch_barcode_num(run) < 2) {  // main.rs:22-25
        fprintf(stderr, "Fewer than 2 counted barcodes.  Too few for barcode enrichment.  Argument flag is ignored\n");
        enrich = false;
    }
    std::vector<bc_ctx*> ctxs;
    auto cleanup = [&]() {
        for (bc_ctx* c : ctxs) bc_destroy(c);
        bch_close(run);
    };
    for (int d : devices) {
        bc_ctx* c = nullptr;
        if (bc_create(bch_config(run), d, 0, &c) != BC_OK) {
            fprintf(stderr, "Error: %s\n", bc_last_error(nullptr));
            cleanup();
            return 1;
        }
        ctxs.push_back(c);
    }
    const bool gz = fastq.size() >= 8 && fastq.compare(fastq.size() - 8, 8, "fastq.gz") == 0;
    if (gz)  // input.rs:60-61
        printf("If this program stops reading before the expected number of sequencing reads, unzip the gzipped fastq and rerun.\n\n");
    // input.rs:54-57, 151-158: the running total, rewritten in place
    bch_set_progress(run, [](uint64_t n, void*) {
        printf("Total sequences:             %s\r", thousands(n).c_str());
        fflush(stdout);
    }, nullptr);
    uint64_t total = 0;
    if (bch_count_fastq_multi(run, ctxs.data(), (int)ctxs.size(), fastq.c_str(), threads, batch_reads, &total, err, sizeof err) != BC_OK) {
        fprintf(stderr, "Error: %s\n", err);
        cleanup();
        return 1;
    }
    // Q21: on the gzip path the reference's read loop feeds one more (empty) line to its line counter at the end of the
    // stream (input.rs:69-73, 129-133), so its total is one above the number of records
    const uint64_t shown_total = total + (gz && total ? 1 : 0);
    uint64_t c[BC_N_COUNTERS] = {0};
    if (bch_counters_multi(ctxs.data(), (int)ctxs.size(), c) != BC_OK) {
        fprintf(stderr, "Error: %s\n", bc_last_error(ctxs[0]));
        cleanup();
        return 1;
    }
    printf("Total sequences:             %s\r\n", thousands(shown_total).c_str());
    printf("Correctly matched sequences: %s\nConstant region mismatches:  %s\nSample barcode mismatches:   %s\n"
           "Counted barcode mismatches:  %s\nDuplicates:                  %s\nLow quality barcodes:        %s\n\n",
           thousands(c[BC_CNT_MATCHED]).c_str(), thousands(c[BC_CNT_CONSTANT]).c_str(), thousands(c[BC_CNT_SAMPLE]).c_str(),
           thousands(c[BC_CNT_COUNTED]).c_str(), thousands(c[BC_CNT_DUPLICATES]).c_str(), thousands(c[BC_CNT_LOW_QUALITY]).c_str());
    if (c[BC_CNT_UNSUPPORTED])
        fprintf(stderr, "WARNING: %s reads hold characters outside A/C/G/T/N (or are longer than %d bases) and were not decoded (the "
                        "reference treats such characters as plain mismatching symbols)\n", thousands(c[BC_CNT_UNSUPPORTED]).c_str(),
                BC_MAX_READ_LEN);
    const auto t1 = std::chrono::steady_clock::now();
    printf("Compute time: %s\n\n-WRITING COUNTS-\n", hms(std::chrono::duration<double>(t1 - t0).count()).c_str());
    static char names[1 << 20];
    const int nfiles = bch_write_counts_multi(run, ctxs.data(), (int)ctxs.size(), outdir.c_str(), prefix.c_str(), merge, enrich, names,
                                              sizeof names, err, sizeof err);
    if (nfiles < 0) {
        fprintf(stderr, "Error: %s\n", err);
        cleanup();
        return 1;
    }
    // File names and "Barcodes counted" as the reference prints them while writing (output.rs:143-165, 355-359, 450-475),
    // then the run record it appends to {prefix}_barcode_stats.txt (output.rs:488-576): same sections; the times are this
    // run's.  (The reference pairs its file list with a count list in which each family's merged count comes first,
    // output.rs:169, 476-479, so with --merge-output its record attributes counts to the wrong files; here every file is
    // listed with its own count.)
    std::vector<std::string> fnames;
    std::vector<unsigned long long> frows;
    for (char* line = strtok(names, "\n"); line; line = strtok(nullptr, "\n")) {
        char* tab = strchr(line, '\t');
        fnames.push_back(tab ? std::string(line, tab) : std::string(line));
        frows.push_back(tab ? strtoull(tab + 1, nullptr, 10) : 0);
    }
    std::string stats_files;
    for (size_t i = 0; i < fnames.size(); i++) {
        const bool merged_file = fnames[i].find("_counts.all") != std::string::npos;
        printf("%s\n", fnames[i].c_str());
        if (merged_file) printf("Barcodes counted: %s\n", thousands(frows[i]).c_str());
        else printf("Barcodes counted: %s\r\n", thousands(frows[i]).c_str());
        stats_files += "File & barcodes counted: " + fnames[i] + "\t" + thousands(frows[i]) + "\n";
    }
    {
        std::string dir = outdir;
        if (!dir.empty() && dir.back() != '/') dir.push_back('/');
        FILE* sf = fopen((dir + prefix + "_barcode_stats.txt").c_str(), "a");
        if (sf) {
            char t_start[32], t_end[32];
            const time_t now = time(nullptr);
            const double elapsed = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            const time_t started = now - (time_t)elapsed;
            strftime(t_start, sizeof t_start, "%Y-%m-%d %H:%M:%S", localtime(&started));
            strftime(t_end, sizeof t_end, "%Y-%m-%d %H:%M:%S", localtime(&now));
            fprintf(sf, "-TIME INFORMATION-\nStart: %s\nFinish: %s\nTotal time: %s\n\n", t_start, t_end, hms(elapsed).c_str());
            fprintf(sf, "-INPUT FILES-\nFastq: %s\nFormat: %s\nSamples: %s\nBarcodes: %s\n\n", fastq.c_str(), format.c_str(),
                    samples.empty() ? "None" : samples.c_str(), counted.empty() ? "None" : counted.c_str());
            fprintf(sf, "%s\n", bch_describe(run));
            fprintf(sf, "-RESULTS-\nTotal sequences:             %s\nCorrectly matched sequences: %s\nConstant region mismatches:  %s\n"
                        "Sample barcode mismatches:   %s\nCounted barcode mismatches:  %s\nDuplicates:                  %s\n"
                        "Low quality barcodes:        %s\n\n",
                    thousands(shown_total).c_str(), thousands(c[BC_CNT_MATCHED]).c_str(), thousands(c[BC_CNT_CONSTANT]).c_str(),
                    thousands(c[BC_CNT_SAMPLE]).c_str(), thousands(c[BC_CNT_COUNTED]).c_str(), thousands(c[BC_CNT_DUPLICATES]).c_str(),
                    thousands(c[BC_CNT_LOW_QUALITY]).c_str());
            fprintf(sf, "-OUTPUT FILES-\n%s\n", stats_files.c_str());
            if (gz && shown_total < 1000000) {  // output.rs:566-571
                const char* warning = "WARNING: The program may have stopped early with the gzipped file.  Unzip the fastq.gz and rerun the "
                                      "algorithm on the unzipped fastq file if the number of reads is expected to be above 1,000,000 ";
                printf("\n%s\n\n", warning);
                fprintf(sf, "\n%s\n", warning);
            }
            fprintf(sf, "--------------------------------------------------------------------------------------------------\n\n\n");
            fclose(sf);
        }
    }
    const auto t2 = std::chrono::steady_clock::now();
    printf("\nTotal time: %s\n", hms(std::chrono::duration<double>(t2 - t0).count()).c_str());
    cleanup();
    return 0;
}
