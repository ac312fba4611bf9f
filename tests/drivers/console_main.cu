/*
 * A console driver written against the reference's Face-2 names only
 * (initialise / singleIteration / cleanup / get_agent_<A>_<S>_count /
 * get_<A>_<S>_variable_<v>, header.xslt:531-645), in the shape of the
 * reference's own console loop (main.xslt:334-370, 405-436): it shows that a
 * driver, step function or visualisation written for the generated
 * simulation.cu links against a model plugin + libfgpu_rt.so unchanged.
 *
 * usage: console <0.xml> <iterations>
 */
#include <cstdio>
#include <cstdlib>

#include "header.h"

int main(int argc, char **argv) {
    if (argc < 3) {
        printf("usage: %s input_path num_iterations\n", argv[0]);
        return 1;
    }
    const int iterations = atoi(argv[2]);
    initialise(argv[1]);
    cudaEvent_t start, stop;
    cudaEventCreate(&start);
    cudaEventCreate(&stop);
    cudaEventRecord(start);
    for (int i = 0; i < iterations; i++) singleIteration();
    cudaEventRecord(stop);
    cudaEventSynchronize(stop);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, start, stop);
    printf("Total Processing time: %f (ms)\n", ms);
    const int n = get_agent_Particle_default_count();
    printf("iteration %u count %d max %d\n", getIterationNumber(), n, get_agent_Particle_MAX_count());
    for (int i = 0; i < n && i < 8; i++)
        printf("agent %d id %d x %.9g y %.9g z %.9g nbr %d\n", i, get_Particle_default_variable_id(i),
               get_Particle_default_variable_x(i), get_Particle_default_variable_y(i), get_Particle_default_variable_z(i),
               get_Particle_default_variable_nbr(i));
    cleanup();
    return 0;
}
