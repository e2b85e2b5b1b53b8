// dto_cli.cpp -- command-line drop-in for the reference binary (src/main.rs:20-166): same flags, the same
// run-information lines on stderr and the same pretty JSON on stdout.  `-t/--threads` is accepted for
// compatibility (the work runs on the GPU); `-m/--multi-node` shards permutations over every visible GPU
// of this box, which is what replaces the MPI path (src/run/multi_node.rs).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <string>
#include <vector>

#include "../../include/dto_b200.h"

static void usage(FILE *f) {
    fprintf(f,
            "Dual Threshold Optimization CLI (B200 build)\n\n"
            "Usage: dual_threshold_optimization [OPTIONS] --ranked-list1 <FILE> --ranked-list2 <FILE>\n\n"
            "Options:\n"
            "  -1, --ranked-list1 <FILE>          Path to the first ranked feature list (CSV: feature,rank; no header)\n"
            "  -2, --ranked-list2 <FILE>          Path to the second ranked feature list (CSV: feature,rank; no header)\n"
            "  -b, --background <FILE>            Path to the background feature list (one feature per line, optional)\n"
            "  -p, --permutations <PERMUTATIONS>  Number of permutations to perform [default: 1000]\n"
            "  -t, --threads <THREADS>            Accepted for compatibility; the work runs on the GPU [default: 1]\n"
            "  -m, --multi-node                   Shard permutations over all visible GPUs (replaces the MPI mode)\n"
            "      --seed <SEED>                  Philox seed of the permutation null [default: drawn from the OS, like the\n"
            "                                     reference's thread_rng; pass a value for a reproducible null]\n"
            "  -h, --help                         Print help\n"
            "  -V, --version                      Print version\n");
}

static int die(const char *what) {
    fprintf(stderr, "error: %s: %s\n", what, dto_b200_last_error());
    return 101;  // rust panic exit code
}

int main(int argc, char **argv) {
    std::string list1, list2, background;
    unsigned long long permutations = 1000, threads = 1, seed = 0;
    bool multi = false, have_seed = false;
    for (int i = 1; i < argc; ++i) {
        std::string a = argv[i];
        auto value = [&](const char *name) -> const char * {
            if (i + 1 >= argc) {
                fprintf(stderr, "error: a value is required for '%s' but none was supplied\n", name);
                exit(2);
            }
            return argv[++i];
        };
        if (a == "-1" || a == "--ranked-list1") list1 = value("--ranked-list1 <FILE>");
        else if (a == "-2" || a == "--ranked-list2") list2 = value("--ranked-list2 <FILE>");
        else if (a == "-b" || a == "--background") background = value("--background <FILE>");
        else if (a == "-p" || a == "--permutations") permutations = strtoull(value("--permutations"), nullptr, 10);
        else if (a == "-t" || a == "--threads") threads = strtoull(value("--threads"), nullptr, 10);
        else if (a == "--seed") {
            seed = strtoull(value("--seed"), nullptr, 10);
            have_seed = true;
        }
        else if (a == "-m" || a == "--multi-node") multi = true;
        else if (a == "-h" || a == "--help") {
            usage(stdout);
            return 0;
        } else if (a == "-V" || a == "--version") {
            printf("dual_threshold_optimization 2.0.1 (%s)\n", dto_b200_version());
            return 0;
        } else {
            fprintf(stderr, "error: unexpected argument '%s' found\n\n", a.c_str());
            usage(stderr);
            return 2;
        }
    }
    if (list1.empty() || list2.empty()) {
        fprintf(stderr, "error: the following required arguments were not provided:\n%s%s\n",
                list1.empty() ? "  --ranked-list1 <FILE>\n" : "", list2.empty() ? "  --ranked-list2 <FILE>\n" : "");
        usage(stderr);
        return 2;
    }
    if (!have_seed) {  // the reference shuffles with thread_rng (permuted.rs:58): a fresh null on every run
        std::random_device rd;
        seed = ((unsigned long long)rd() << 32) ^ (unsigned long long)rd();
    }
    if (threads == 0) {
        fprintf(stderr, "Warning: Number of threads cannot be 0. Setting threads to 1.\n");
        threads = 1;
    }
    fprintf(stderr, "Ranked list 1: %s\n", list1.c_str());
    fprintf(stderr, "Ranked list 2: %s\n", list2.c_str());
    fprintf(stderr, "Permutations: %llu\n", permutations);
    fprintf(stderr, "Threads: %llu\n", threads);
    fprintf(stderr, "Multi-node mode: %s\n", multi ? "enabled" : "disabled");

    dto_b200_ranked_list *l1 = nullptr, *l2 = nullptr;
    dto_b200_feature_list *bg = nullptr;
    if (dto_b200_read_ranked_list_csv(list1.c_str(), &l1)) return die("ranked list 1");
    if (dto_b200_read_ranked_list_csv(list2.c_str(), &l2)) return die("ranked list 2");
    fprintf(stderr,
            "The product of the lengths of the threshold lists (this describes the asymptotic runtime of a single job): "
            "%zu\n",
            dto_b200_ranked_list_num_thresholds(l1) * dto_b200_ranked_list_num_thresholds(l2));
    if (!background.empty() && dto_b200_read_feature_list(background.c_str(), &bg)) return die("background");
    uint64_t population = 0;
    if (dto_b200_compute_population_size(l1, l2, bg, &population)) return die("compute_population_size");

    // tasks = [Task{0, permute:false}] + P x Task{id, permute:true}   (main.rs:122-126)
    std::vector<uint8_t> task_permute((size_t)permutations + 1, 1);
    task_permute[0] = 0;
    std::vector<dto_b200_record> records(task_permute.size());
    std::vector<int> devices;
    if (multi) {
        int n = 0;
        if (dto_b200_device_count(&n)) return die("device count");
        for (int d = 0; d < n; ++d) devices.push_back(d);
        fprintf(stderr, "GPUs: %d\n", n);
    }
    if (dto_b200_run_single_node(l1, l2, population, task_permute.data(), task_permute.size(), devices.data(),
                                 devices.size(), seed, records.data()))
        return die("run");
    // the reference's tie notices (optimize_main.rs:85-107) for the unpermuted task
    if (records[0].flags & DTO_B200_FLAG_TIE_MINP) {
        fprintf(stderr, "Multiple results with the same minimum p-value (%.15f). Choosing the result with the largest intersection size.\n",
                records[0].pvalue);
        if (records[0].flags & DTO_B200_FLAG_TIE_OVERLAP)
            fprintf(stderr, "Multiple results with the same maximum intersection size (%u). Choosing an arbitrary result based on the order of thresholds.\n",
                    records[0].intersection_size);
    }
    dto_b200_final_result fin;
    if (dto_b200_empirical_pvalue(records.data(), records.size(), &fin)) return die("empirical_pvalue");
    size_t len = 0;
    dto_b200_final_result_json(&fin, nullptr, 0, &len);
    std::string json(len + 1, '\0');
    dto_b200_final_result_json(&fin, json.data(), json.size(), &len);
    printf("%s\n", json.c_str());
    dto_b200_ranked_list_free(l1);
    dto_b200_ranked_list_free(l2);
    dto_b200_feature_list_free(bg);
    return 0;
}
