// oracle/oracle_cli.cpp — `barcode-count-oracle`: the oracle behind the reference's flag surface
// (arguments.rs:27-124).  TEST INFRASTRUCTURE ONLY: used as the CPU baseline and for CSV parity runs.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <string>
#include <thread>

#include "oracle.hpp"

static void usage() {
    fprintf(stderr,
            "barcode-count-oracle (CPU restatement of NGS-Barcode-Count 0.11.1; test infrastructure)\n"
            "  -f, --fastq <file>              FastQ file (.fastq / .fastq.gz)\n"
            "  -q, --sequence-format <file>    Sequence format file\n"
            "  -s, --sample-barcodes <file>    Sample barcodes file\n"
            "  -c, --counted-barcodes <file>   Counted barcodes file\n"
            "  -t, --threads <n>               Number of threads\n"
            "  -o, --output-dir <dir>          Directory to output the counts to [./]\n"
            "  -p, --prefix <str>              File prefix name [today]\n"
            "  -m, --merge-output              Merge sample output counts into a single file\n"
            "  -e, --enrich                    Single/double barcode enrichment files\n"
            "      --max-errors-counted-barcode <n>\n"
            "      --max-errors-sample <n>\n"
            "      --max-errors-constant <n>\n"
            "      --min-quality <f>           Minimum average read quality score per barcode [0]\n");
}

int main(int argc, char** argv) {
    std::string fastq, format, samples, counted;
    oracle::Options opt;
    unsigned threads = std::thread::hardware_concurrency();
    {
        char buf[32];
        time_t now = time(nullptr);
        strftime(buf, sizeof buf, "%Y-%m-%d", localtime(&now));  // arguments.rs:25
        opt.prefix = buf;
    }
    for (int i = 1; i < argc; i++) {
        std::string a = argv[i];
        auto val = [&]() -> std::string {
            if (i + 1 >= argc) {
                fprintf(stderr, "error: %s needs a value\n", a.c_str());
                exit(2);
            }
            return argv[++i];
        };
        if (a == "-f" || a == "--fastq") fastq = val();
        else if (a == "-q" || a == "--sequence-format") format = val();
        else if (a == "-s" || a == "--sample-barcodes") samples = val();
        else if (a == "-c" || a == "--counted-barcodes") counted = val();
        else if (a == "-t" || a == "--threads") threads = static_cast<unsigned>(std::stoul(val()));
        else if (a == "-o" || a == "--output-dir") opt.output_dir = val();
        else if (a == "-p" || a == "--prefix") opt.prefix = val();
        else if (a == "-m" || a == "--merge-output") opt.merge_output = true;
        else if (a == "-e" || a == "--enrich") opt.enrich = true;
        else if (a == "--max-errors-counted-barcode") opt.barcodes_errors = static_cast<uint16_t>(std::stoul(val()));
        else if (a == "--max-errors-sample") opt.sample_errors = static_cast<uint16_t>(std::stoul(val()));
        else if (a == "--max-errors-constant") opt.constant_errors = static_cast<uint16_t>(std::stoul(val()));
        else if (a == "--min-quality") opt.min_quality = std::stof(val());
        else if (a == "-h" || a == "--help") { usage(); return 0; }
        else { fprintf(stderr, "error: unknown argument %s\n", a.c_str()); usage(); return 2; }
    }
    if (fastq.empty() || format.empty()) {
        usage();
        return 2;
    }
    try {
        auto t0 = std::chrono::steady_clock::now();
        oracle::Pipeline p(format, samples, counted, opt);
        printf("%s\n\n%s\n\n", p.format.display().c_str(), p.max_errors.display().c_str());
        unsigned long long total = p.run_fastq(fastq, threads);
        auto t1 = std::chrono::steady_clock::now();
        printf("Total sequences:             %llu\n%s\n\n", total, p.errors.display().c_str());
        double compute = std::chrono::duration<double>(t1 - t0).count();
        printf("Compute time: %.3f seconds (%u threads, %.0f reads/s)\n\n-WRITING COUNTS-\n", compute, threads,
               compute > 0 ? static_cast<double>(total) / compute : 0.0);
        for (const std::string& f : p.write_counts_files()) printf("%s\n", f.c_str());
        auto t2 = std::chrono::steady_clock::now();
        printf("\nTotal time: %.3f seconds\n", std::chrono::duration<double>(t2 - t0).count());
    } catch (const std::exception& e) {
        fprintf(stderr, "Error: %s\n", e.what());
        return 1;
    }
    return 0;
}
