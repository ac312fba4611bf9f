/*
 * fgpu_console -- the reference's console mode (main.xslt:60-448) for any model plugin, over the C ABI.
 *
 *   fgpu_console model_plugin.so input_path num_iterations [cuda_device_id] [XML_output_frequency]
 *
 * Same arguments after the plugin path, same defaults and messages as the generated main.cu:
 * input_path is the initial-states XML file (its directory receives the iteration files) or an output
 * directory; num_iterations >= 0 (0 = until a step function calls set_exit_early); XML output every
 * XML_output_frequency iterations (default 1, 0 = none) plus the final iteration; "Total Processing time"
 * around the iteration loop.  The reference builds one such executable per model; here the model is a
 * plugin, so one driver serves all of them.
 */
#include <sys/stat.h>

#include <cerrno>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#include "fgpu.h"

static const int OUTPUT_TO_XML = 1; /* main.xslt:29 */

static void die() {
    fprintf(stderr, "%s\n", fgpu_last_error());
    fflush(stderr);
    exit(EXIT_FAILURE);
}

static bool path_properties(const std::string &path, bool *is_file, bool *is_dir) { /* main.xslt:173-205 */
    struct stat st;
    *is_file = *is_dir = false;
    if (stat(path.c_str(), &st) == 0) {
        *is_dir = (st.st_mode & S_IFDIR) != 0;
        *is_file = (st.st_mode & S_IFREG) != 0;
        return *is_dir || *is_file;
    }
    if (errno == ENOENT) return false;
    fprintf(stderr, "Error: An unknown error occured while processing file infomration.\n");
    exit(EXIT_FAILURE);
}

static std::string parent_directory(const std::string &path) { /* main.xslt:150-166: keeps the trailing slash */
    const size_t k = path.find_last_of("/\\");
    return k == std::string::npos ? std::string() : path.substr(0, k + 1);
}

int main(int argc, char **argv) {
    bool help = false;
    for (int i = 1; i < argc; ++i) help = help || !strcmp(argv[i], "-h") || !strcmp(argv[i], "--help");
    printf("FLAMEGPU Console mode\n");
    if (help || argc < 4 || argc > 6) {
        printf("\nusage: %s [-h] [--help] model_plugin input_path num_iterations [cuda_device_id] [XML_output_override]\n", argv[0]);
        printf("\nrequired arguments:\n");
        printf("  model_plugin         Path to the model plugin (.so) built from XMLModelFile.xml + functions.c\n");
        printf("  input_path           Path to initial states XML file OR path to output XML directory\n");
        printf("  num_iterations       Number of simulation iterations\n");
        printf("\noptions arguments:\n");
        printf("  -h, --help           Output this help message.\n");
        printf("  cuda_device_id       CUDA device ID to be used. Default is 0.\n");
        printf("  XML_output_frequency Frequency of XML output\n");
        printf("                         0 = No output\n                         1 = Every 1 iteration\n                         5 = Every 5 iterations\n");
        printf("                         Default value: %d\n", OUTPUT_TO_XML);
        return EXIT_FAILURE;
    }
    /* setFilePaths, main.xslt:207-254 */
    std::string input = argv[2], inputfile, outputpath;
    bool is_file = false, is_dir = false;
    if (path_properties(input, &is_file, &is_dir)) {
        if (is_file) {
            inputfile = input;
            outputpath = parent_directory(input);
        } else
            outputpath = input;
    } else {
        outputpath = parent_directory(input);
        bool f = false, d = false;
        if (!outputpath.empty() && path_properties(outputpath, &f, &d)) {
            if (!d || f) {
                printf("Error: outputpath `%s` exists, but it is not a directory.\n", outputpath.c_str());
                return EXIT_FAILURE;
            }
            printf("Warning: `%s` does not exist using parent directory for output.\n", input.c_str());
        } else {
            printf("Warning: Parent directory `%s` does not exist. Using current working directory for output.\n", outputpath.c_str());
            outputpath.clear();
        }
    }
    int frequency = OUTPUT_TO_XML; /* getOutputXMLFrequency, main.xslt:256-274 */
    if (argc >= 6) frequency = atoi(argv[5]) <= 0 ? 0 : atoi(argv[5]);
    const int device = argc >= 5 ? atoi(argv[4]) : 0;
    const int iterations = atoi(argv[3]);
    if (iterations < 0) {
        printf("Second argument must be a positive integer (Number of Iterations), or 0 for inifinite iterations\n");
        return EXIT_FAILURE;
    }

    const fgpu_model *model = fgpu_model_load(argv[1]);
    if (!model) die();
    fgpu_sim *sim = fgpu_sim_create(model, device);
    if (!sim) die();
    fgpu_set_output_dir(outputpath.c_str()); /* getOutputDir() of step functions */
    /* The reference's -DINSTRUMENT_ITERATIONS / _AGENT_FUNCTIONS / _INIT_ / _STEP_ / _EXIT_FUNCTIONS are compile-time
     * (simulation.xslt:370-377); a plugin is not recompiled for them, so the console takes the same switches from the
     * environment: FGPU_INSTRUMENT=iterations,agent_functions,... */
    if (const char *ins = getenv("FGPU_INSTRUMENT")) {
        const std::string want(ins);
        for (const char *k : {"iterations", "agent_functions", "init_functions", "step_functions", "exit_functions"})
            if (want.find(k) != std::string::npos && fgpu_sim_set_option(sim, (std::string("instrument_") + k).c_str(), 1) != FGPU_OK) die();
    }
    if (!inputfile.empty()) {
        printf("Initial states: %s\n", inputfile.c_str());
        if (fgpu_sim_load_states(sim, inputfile.c_str()) != FGPU_OK) die();
    }
    if (fgpu_sim_run_init_functions(sim) != FGPU_OK) die();

    const auto t0 = std::chrono::steady_clock::now();
    int done = 0;
    for (int i = 0; i < iterations || iterations == 0; ++i) {
        printf("Processing Simulation Step %i\n", i + 1);
        if (fgpu_sim_step(sim, 1) != FGPU_OK) die();
        done = i + 1;
        if (frequency > 0 && (i + 1) % frequency == 0) {
            if (fgpu_sim_save_states(sim, outputpath.c_str(), i + 1) != FGPU_OK) die();
            printf("Iteration %i Saved to XML\n", i + 1);
        }
        if (fgpu_get_exit_early()) break;
    }
    if (frequency > 0 && done % frequency != 0) {
        if (fgpu_sim_save_states(sim, outputpath.c_str(), done) != FGPU_OK) die();
        printf("Iteration %i Saved to XML\n", done);
    }
    const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    printf("Total Processing time: %f (ms)\n", ms);
    if (fgpu_sim_run_exit_functions(sim) != FGPU_OK) die();
    fgpu_sim_destroy(sim);
    return EXIT_SUCCESS;
}
