// hmc -- the HMC command line (HMC.cpp / Main.cpp) on the B200 path.  Same option names and defaults as the
// reference's boost::program_options table (HMC.cpp:19-56), parsed by hand (Boost is not a dependency);
// resolve mode only: reads a PHASE file and writes <file>.reconstructed (HMC.cpp:198-233).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "hmc_host.hpp"

using namespace hmc_host;

static void usage()
{
    printf("Usage: hmc [option ...] datafile\n"
           "Generic options:\n"
           "  -h [ --help ]                 Show help message\n"
           "Configurations:\n"
           "  --nologo                      Suppress logo and copyright information\n"
           "  -d [ --debug ] arg (=4)       Set debug level\n"
           "  -f [ --input-format ] arg (=PHASE)  Set input file format (PHASE only)\n"
           "Model parameters:\n"
           "  -m [ --model ] arg (=MV)      Set the inference model (MV|MC|MA)\n"
           "  -r [ --min-freq-rel ] arg     Minimum relative frequency of patterns\n"
           "  -a [ --min-freq-abs ] arg (=1.5)  Minimum absolute frequency of patterns\n"
           "  -n [ --num-patterns ] arg     Maximum number of patterns\n"
           "  --min-pattern-len arg (=1)    Minimum length of patterns\n"
           "  --max-pattern-len arg (=30)   Maximum length of patterns\n"
           "  -o [ --mc-order ] arg (=1)    Markov chain order\n"
           "  --exact-estimate              Re-estimate frequency exactly (i.e. not using sampling)\n"
           "  --sample-size arg (=10)       Sample some most probable configurations\n"
           "  --max-sample-size arg         Maximum sample size (accepted, unused -- as in the reference)\n"
           "  --final-sample-size arg (=1)  Final sample size (accepted, unused)\n"
           "  -i [ --max-iteration ] arg (=1)  Maximum iteration number\n"
           "  --min-iteration arg (=-1)     Minimum iteration number (accepted, unused)\n"
           "Device options (new):\n"
           "  --to-binary arg               Convert the input panel to the bit-packed binary format (HMCB) and exit\n"
           "  -e [ --compare ]              Compare input data with target data: hmc -e INFERRED.phase TARGET.phase\n"
           "  --device arg (=0)             CUDA device\n"
           "  --scaled                      carry a power-of-two exponent so that long panels do not underflow\n"
           "  --threads arg (=0)            threads per individual on the device (0 = auto)\n");
}

int main(int argc, char **argv)
{
    try {
        std::string model = "MV", input, target, to_binary, format = "PHASE";
        bool compare = false;
        int device = 0, debug = 4, threads = 0;
        bool nologo = false, scaled = false;
        // first pass for device options
        for (int i = 1; i < argc; ++i) {
            if (!strcmp(argv[i], "--device") && i + 1 < argc) device = atoi(argv[i + 1]);
            if (!strcmp(argv[i], "--threads") && i + 1 < argc) threads = atoi(argv[i + 1]);
        }
        std::vector<std::string> args(argv + 1, argv + argc);
        bool help = args.empty();
        for (const std::string &a : args) if (a == "-h" || a == "--help") help = true;
        if (help) { usage(); return 0; }
        HaploModel hm(device, 0, threads);
        hm.setCliDefaults();
        for (size_t i = 0; i < args.size(); ++i) {
            const std::string &a = args[i];
            auto next = [&]() -> std::string { if (i + 1 >= args.size()) throw HmcError("option " + a + " needs a value"); return args[++i]; };
            if (a == "--nologo") nologo = true;
            else if (a == "-d" || a == "--debug") debug = atoi(next().c_str());
            else if (a == "-f" || a == "--input-format") format = next();
            else if (a == "-m" || a == "--model") model = next();
            else if (a == "-r" || a == "--min-freq-rel") hm.min_freq = atof(next().c_str());
            else if (a == "-a" || a == "--min-freq-abs") hm.min_freq_abs = atof(next().c_str());
            else if (a == "-n" || a == "--num-patterns") hm.num_patterns = atoi(next().c_str());
            else if (a == "--min-pattern-len") hm.min_pattern_len = atoi(next().c_str());
            else if (a == "--max-pattern-len") hm.max_pattern_len = atoi(next().c_str());
            else if (a == "-o" || a == "--mc-order") hm.mc_order = atoi(next().c_str());
            else if (a == "--exact-estimate") hm.exact_estimate = true;
            else if (a == "--sample-size") hm.sample_size = atoi(next().c_str());
            else if (a == "--max-sample-size") hm.max_sample_size = atoi(next().c_str());
            else if (a == "--final-sample-size") hm.final_sample_size = atoi(next().c_str());
            else if (a == "-i" || a == "--max-iteration") hm.max_iteration = atoi(next().c_str());
            else if (a == "--min-iteration") hm.min_iteration = atoi(next().c_str());
            else if (a == "--device" || a == "--threads") next();
            else if (a == "--scaled") scaled = true;
            else if (a == "-e" || a == "--compare") compare = true;
            else if (a == "--to-binary") to_binary = next();
            else if (!a.empty() && a[0] == '-') throw HmcError("unknown option " + a);
            else if (input.empty()) input = a;
            else target = a;
        }
        if (format != "PHASE") throw HmcError("Unknown input format " + format + "! (this build reads PHASE files)");
        if (input.empty()) { usage(); return 0; }
        if (!nologo && debug >= 3) {
            printf("HMC - Haplotype inference tool based on Markov Chain [hmc_b200, %s]\n", hmc_version());
        }
        hm.setModel(model);
        hm.setScaled(scaled);
        hm.verbose = debug >= 4;
        HaploFile file(input);
        GenoData genos;
        if (debug >= 3) printf("Reading genotype file ...\n");
        file.readGenoData(genos);
        if (debug >= 3) printf("Succesfully read Haplotype file with %d markers and %d genotypes.\n", genos.L, genos.N);
        if (!to_binary.empty()) {   // PHASE text -> binary panel (same data, bit-packed; read back by every mode of this tool)
            HaploFile::writeBinary(to_binary, genos, genos.idx);
            if (debug >= 3) printf("Wrote binary panel %s\n", to_binary.c_str());
            return 0;
        }
        if (compare) {       // HMC::compare (HMC.cpp:246-257): the first file is the inferred data, the second the target
            if (target.empty()) throw HmcError("Input format " + format + " require 1 filenames! (--compare needs INPUT and TARGET)");
            GenoData tg;
            HaploFile(target).readGenoData(tg);
            if (debug >= 3) printf("Succesfully read Haplotype file with %d markers and %d genotypes.\n", tg.L, tg.N);
            HaploComp hc(hm, tg, genos);
            printf("\nSwitch Error = %f, IHP = %f, IGP = %f\n", hc.switch_error(), hc.incorrect_haplotype_percentage(), hc.incorrect_genotype_percentage());
            return 0;
        }
        std::vector<int8_t> resolutions;
        hm.run(genos, resolutions);
        if (debug >= 3) {
            for (size_t i = 0; i < hm.ll_trace.size(); ++i) printf("  iteration %zu: patterns = %d, LL = %f\n", i + 1, hm.npat_trace[i], hm.ll_trace[i]);
            printf("Solving Time = %f (mining %.3f s, device %.3f s, host M-step %.3f s)\n", hm.seconds[0] + hm.seconds[1] + hm.seconds[2], hm.seconds[0], hm.seconds[1], hm.seconds[2]);
        }
        if (HaploFile::isBinary(input)) HaploFile::writeBinary(input + ".reconstructed", genos, resolutions);     // output in the input's format
        else hm.writeReconstructed(file, genos, resolutions, ".reconstructed");
        return 0;
    }
    catch (const std::exception &e) {
        fprintf(stderr, "ERROR: %s\n", e.what());
        return 1;
    }
}
