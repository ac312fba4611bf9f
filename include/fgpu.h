/*
 * fgpu.h -- C ABI of the B200-native FLAME GPU 1.5 simulation-step runtime
 * (libfgpu_rt.so).
 *
 * The reference has no FFI: its seam is source level.  Each
 * entry point below names the generated reference function it replaces
 * (line numbers are lines of the .xslt templates under
 * /root/reference/FLAMEGPU/templates/).  All functions return 0 on success
 * and a negative code on failure (the reference prints and calls exit(),
 * simulation.xslt:198-206); `fgpu_last_error()` returns the message.
 *
 * Threading: one fgpu_sim = one host thread + its CUDA stream.  Objects are
 * independent; the core has no global state.
 * Ownership: every pointer passed in is borrowed for the call.  Host buffers
 * handed to read/write calls must hold `count` elements of the variable type.
 */
#ifndef FGPU_H
#define FGPU_H

#include <stddef.h>
#include <stdint.h>
#include "fgpu_model.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct fgpu_sim fgpu_sim;

enum {
    FGPU_OK = 0,
    FGPU_ERR_ARG = -1,        /* bad index / null pointer */
    FGPU_ERR_CUDA = -2,       /* a CUDA call failed (gpuErrchk, simulation.xslt:198) */
    FGPU_ERR_OVERFLOW = -3,   /* a buffer-size guard fired (simulation.xslt:1609,1816,1940; io.xslt:406) */
    FGPU_ERR_IO = -4,         /* file could not be read */
    FGPU_ERR_MODEL = -5,      /* plugin missing / ABI mismatch */
    FGPU_ERR_COMM = -6        /* slab exchange failed */
};

const char *fgpu_last_error(void);

/* ---- model plugins ------------------------------------------------------ */
/* dlopen()s a model plugin built from XMLModelFile.xml + functions.c and
 * returns its descriptor (replaces the xsltproc + nvcc step, tools/common.mk:347-451). */
const fgpu_model *fgpu_model_load(const char *plugin_path);

/* ---- lifecycle ----------------------------------------------------------- */
/* initialise() without the XML read: allocation, partition constants, rand48
 * seeding (simulation.xslt:472-747).  `device` is the CUDA ordinal (main.xslt:311). */
fgpu_sim *fgpu_sim_create(const fgpu_model *model, int device);
/* cleanup() (simulation.xslt:779-876). */
void fgpu_sim_destroy(fgpu_sim *sim);

/* readInitialStates() (io.xslt:211-621): parse a 0.xml into the initial state
 * lists, apply <environment> constants, upload. */
int fgpu_sim_load_states(fgpu_sim *sim, const char *xml_path);
/* The same reader without a device (host code only): how many agents of type `agent` a states file holds -- e.g. to
 * choose bufferSize before generating a model -- and variable `var` of the first `count` of them, in file order.
 * <environment> values are not applied.  Returns the count / FGPU_OK, or a negative error code. */
int fgpu_states_file_count(const fgpu_model *model, const char *xml_path, int agent);
int fgpu_states_file_read_var(const fgpu_model *model, const char *xml_path, int agent, int var, void *dst, int count);

/* Process-wide driver state of the reference's main.cu that step functions may use: getOutputDir()
 * (main.xslt:139-141: where iteration files and custom outputs go) and set_exit_early()/get_exit_early()
 * (header.xslt:569-576: a step function asks the driver loop to stop).  The plugin shim forwards to these. */
void fgpu_set_output_dir(const char *path);
const char *fgpu_get_output_dir(void);
void fgpu_set_exit_early(int flag);
int fgpu_get_exit_early(void);

/* saveIterationData() (io.xslt:124-200): writes <outputpath><iteration_number>.xml in the reference's
 * format -- <states><itno/><environment/> then one <xagent> per agent, per agent type and state in model
 * order, every value through the reference's format specifier (%f for float/double: lossy, like the
 * reference).  A file written here is readable by fgpu_sim_load_states and by the reference's reader. */
int fgpu_sim_save_states(fgpu_sim *sim, const char *outputpath, int iteration_number);
/* Exact checkpoint / resume (the XML route loses float bits, the rand48 state and the iteration number):
 * a little-endian binary image of every state list, the rand48 seeds, the environment constants and the
 * iteration counter.  Loading checks the model name and the variable layout. */
int fgpu_sim_save_snapshot(fgpu_sim *sim, const char *path);
int fgpu_sim_load_snapshot(fgpu_sim *sim, const char *path);

/* Direct upload of one state list from host SoA columns (one pointer per
 * agent variable, XML order; NULL = defaultValue).  Replaces the malloc +
 * cudaMemcpy H2D of the whole list struct (simulation.xslt:568-581).
 * An agent array variable (arrayLength L, header.xslt:187-194) is one column of L blocks of `count` values,
 * element-major: element e of agent i at column[e * count + i] -- the order of the list on the device with
 * `count` in place of xmachine_memory_<A>_MAX.  Every call that moves columns uses this layout. */
int fgpu_sim_set_agents(fgpu_sim *sim, int agent, int state, int count, const void *const *columns);

/* The same upload without the final synchronisation: the copies are enqueued on the simulation stream and the call
 * returns (columns should come from pinned host memory that stays untouched until fgpu_sim_sync, for the copies to
 * be asynchronous; a NULL column is filled with its default value by a kernel, no traffic over PCIe).  With fgpu_sim_step_async and fgpu_sim_read_agents_async, two simulations can be
 * driven in alternation so that one's upload and read-back overlap the other's iteration (both PCIe directions and
 * the SMs busy at once); the reference's initialise / singleIteration / get_host_* sequence is strictly serial. */
int fgpu_sim_set_agents_async(fgpu_sim *sim, int agent, int state, int count, const void *const *columns);

/* sort_<A>s_<S>() (simulation.xslt:750-776): `generate_key_value_pairs` is the caller's __global__ kernel
 * (unsigned int* keys, unsigned int* values, xmachine_memory_<A>_list* agents), launched over the state list; the
 * (key, value) pairs are sorted by key with the runtime's own LSD radix sort (stable, where the reference's
 * thrust::sort_by_key promises no order among equal keys) and agent i of the new list is agent values[i] of the old. */
int fgpu_sim_sort_agents(fgpu_sim *sim, int agent, int state, const void *generate_key_value_pairs);

/* Host-side mirror of h_add_agents_<A>_<S> (simulation.xslt:1275-1320): append. */
int fgpu_sim_add_agents(fgpu_sim *sim, int agent, int state, int count, const void *const *columns);

/* Init / step / exit functions of the model are run by fgpu_sim_run_init_functions
 * (simulation.xslt:703-716) and inside fgpu_sim_step (simulation.xslt:933-946). */
int fgpu_sim_run_init_functions(fgpu_sim *sim);
int fgpu_sim_run_exit_functions(fgpu_sim *sim);

/* ---- stepping ------------------------------------------------------------ */
/* n x singleIteration() (simulation.xslt:878-975).  Models whose functions have
 * no condition / death / birth run as a captured CUDA graph with no host
 * round trips; the call returns after the work is enqueued AND complete. */
int fgpu_sim_step(fgpu_sim *sim, int n);
/* Same, but returns as soon as the work is enqueued on the simulation stream. */
int fgpu_sim_step_async(fgpu_sim *sim, int n);
int fgpu_sim_sync(fgpu_sim *sim);
/* n iterations timed one by one with cudaEvent pairs on the simulation stream
 * (INSTRUMENT_ITERATIONS, simulation.xslt:881-883,969-974).  flush_bytes > 0
 * rewrites a scratch buffer of that size before each iteration, outside the
 * timed span, so every iteration starts from a cold L2. */
int fgpu_sim_step_timed(fgpu_sim *sim, int n, size_t flush_bytes, float *ms_each);

/* Fine-grained stepping used by parity tests: the prologue of singleIteration
 * (message counts := 0, simulation.xslt:885-892) and one <A>_<f>(stream)
 * wrapper (simulation.xslt:1402-1974). */
int fgpu_sim_begin_iteration(fgpu_sim *sim);
int fgpu_sim_run_function(fgpu_sim *sim, int function);
unsigned int fgpu_sim_iteration(const fgpu_sim *sim);      /* getIterationNumber() */

/* ---- queries (get_agent_<A>_<S>_count etc., header.xslt:588-645) --------- */
int fgpu_sim_agent_count(fgpu_sim *sim, int agent, int state);
/* Copies variable `var` of the first `count` agents of a state list to `dst`. */
int fgpu_sim_read_agent_var(fgpu_sim *sim, int agent, int state, int var, void *dst, int count);
/* The read-back counterpart of fgpu_sim_set_agents: columns[v] receives variable v of the first `count`
 * agents of the state list, NULL entries are skipped.  All copies are issued before one synchronisation,
 * so pinned destinations are filled at bus speed. */
int fgpu_sim_read_agents(fgpu_sim *sim, int agent, int state, int count, void *const *columns);
/* ... enqueued only: the destination arrays are complete after the next fgpu_sim_sync */
int fgpu_sim_read_agents_async(fgpu_sim *sim, int agent, int state, int count, void *const *columns);
/* get_<A>_<S>_variable_<v>(index) (header.xslt:645): one element of a state list. */
int fgpu_sim_read_agent_value(fgpu_sim *sim, int agent, int state, int var, unsigned int index, void *dst);
/* get_<A>_<S>_variable_<v>(index, element) (header.xslt:648-658, simulation.xslt:1093-1140): one element of an
 * agent array variable. */
int fgpu_sim_read_agent_element(fgpu_sim *sim, int agent, int state, int var, unsigned int index, unsigned int element, void *dst);
/* reduce_/min_/max_/count_<A>_<S>_<v>_variable() (header.xslt:729-751, simulation.xslt:1325-1357): one scalar
 * agent variable of a state list folded on the device.  `result` (and `count_value`, FGPU_REDUCE_COUNT only)
 * have the variable's own C type; the count is returned in that type too, like the reference.  Integer results
 * are exact; a float SUM is a fixed-order tree (the reference's thrust::reduce leaves the order open).
 * MIN/MAX of an empty list is an error. */
enum { FGPU_REDUCE_SUM = 0, FGPU_REDUCE_MIN = 1, FGPU_REDUCE_MAX = 2, FGPU_REDUCE_COUNT = 3 };
int fgpu_sim_reduce_agent_var(fgpu_sim *sim, int agent, int state, int var, int op, const void *count_value, void *result);
/* get_device_<A>_<S>_agents(): borrowed device pointer to a list in the
 * reference's xmachine_memory_<A>_list layout; invalidated by the next step. */
void *fgpu_sim_device_agents(fgpu_sim *sim, int agent, int state);

int fgpu_sim_message_count(fgpu_sim *sim, int message);
/* Sorted message list, variable `var`, unpacked to a dense host array. */
int fgpu_sim_read_message_var(fgpu_sim *sim, int message, int var, void *dst, int count);
/* Partition boundary matrix in normalised form: for each cell,
 * count[c] and start[c] (start is only meaningful where count > 0). */
int fgpu_sim_read_pbm(fgpu_sim *sim, int message, int *start, int *count);
/* Cell key of each *unsorted* message (message_<M>_hash, kernals.xslt:877-889)
 * and the sorted permutation (value array of the radix sort, sim:1867). */
int fgpu_sim_read_message_keys(fgpu_sim *sim, int message, unsigned int *keys, int count);
int fgpu_sim_read_message_order(fgpu_sim *sim, int message, unsigned int *order, int count);
/* rand48 seeds, 2 x uint32 per slot (simulation.xslt:681-686). */
int fgpu_sim_read_seeds(fgpu_sim *sim, unsigned int *dst_xy, int count);
int fgpu_sim_rand48_constants(fgpu_sim *sim, unsigned int *AxAyCxCy);

/* set_<constant>() by name (header.xslt:771). */
int fgpu_sim_set_constant(fgpu_sim *sim, const char *name, const void *value);
int fgpu_sim_get_constant(fgpu_sim *sim, const char *name, void *value);

/* ---- instrumentation (INSTRUMENT_ITERATIONS, simulation.xslt:370-377) ---- */
/* Kernel launches issued by the simulation stream since creation. */
long long fgpu_sim_launch_count(const fgpu_sim *sim);
/* Device milliseconds of the last fgpu_sim_step call (cudaEvent pair on the
 * simulation stream, like main.xslt:405-436). */
float fgpu_sim_last_step_ms(const fgpu_sim *sim);
/* Options (int values).  Behaviour: "instrument_iterations", "instrument_agent_functions", "instrument_init_functions",
 * "instrument_step_functions", "instrument_exit_functions" (0/1: the reference's -DINSTRUMENT_* output lines,
 * simulation.xslt:370-377,714,796,913,945,973), "nvtx" (0/1: every stage of an iteration is an NVTX range on the
 * enqueueing thread, like PROFILE_SCOPED_RANGE), "xml_float_digits" (0 = the reference's "%f"; n = "%.<n>g").
 * Tuning, never changing a result: "order" (0/1 cell-coherent processing of message consumers), "graph" (0/1 CUDA-graph
 * replay), "pdl" (0/1 programmatic dependent launch), "radix_bits" (4..11), "msd" (0/1 MSD message build), "carry" (0/1
 * carried agent records), "mirror" (0/1: a cell-order consumer takes the variables its own message carries from that
 * message, fgpu_launch.mirror; on by default), "l2_fetch" (32/64/128).  Slab mode, before fgpu_sim_slab_init: "slab_p2p", "slab_async";
 * any time: "slab_inplace" (0 = stable re-compaction at migration: another list order), "slab_overlap", "slab_direct" (1 = migrants
 * are packed straight into the neighbours' arrival buffers over NVLink, 0 = into a local buffer that a second kernel pushes; 0 by default, measured faster). */
int fgpu_sim_set_option(fgpu_sim *sim, const char *key, int value);
/* Per-kernel-class device time of the most recent profiled iteration.
 * Writes up to `cap` (name,ms) pairs; returns the number written. */
int fgpu_sim_profile_iteration(fgpu_sim *sim, const char **names, float *ms, int cap);
/* Same over `iters` back-to-back iterations (no host synchronisation in between, pooled events):
 * steady-state device time per stage, averaged by stage name. */
int fgpu_sim_profile_iterations(fgpu_sim *sim, int iters, const char **names, float *ms, int cap);

/* ---- kernel-level entry points for the parity tests ---------------------- */
/* The hand-written LSD radix sort (replaces cub::DeviceRadixSort::SortPairs,
 * simulation.xslt:1867): sorts (keys[i], i) on the low `key_bits` bits with
 * digits of at most `max_digit_bits` bits; host in, host out. */
int fgpu_debug_sort_pairs(const unsigned int *keys, int n, int key_bits, int max_digit_bits,
                          unsigned int *keys_out, unsigned int *vals_out);
/* device milliseconds of one such sort of n pseudo-random keys (average of reps): a tuning aid */
int fgpu_debug_sort_time(int n, int key_bits, int max_digit_bits, int reps, float *ms_out);
/* Stable compaction of payload[i] where flags[i] == 1 (scatter_<A>_Agents, kernals.xslt:239-257). */
int fgpu_debug_compact(const int *flags, const unsigned int *payload, int n, unsigned int *out, int *count_out);

/* ---- 1-D slab decomposition over the GPUs of one node ------ */
/* The reference is single-GPU (one cudaSetDevice, main.xslt:311); this is the one
 * parallelism strategy the B200 runtime adds.  The partition grid is cut into
 * `world` slabs of whole cell planes along the slowest hash axis (z; y for 2-D
 * grids, kernals.xslt:888).  After every spatial message build the two boundary
 * planes of the sorted message list travel to the neighbouring ranks (including
 * the wrap pair 0 <-> world-1, because message_<M>_hash sends out-of-range
 * neighbour cells to the opposite face, kernals.xslt:880-885); after every
 * function that moves agents, agents whose plane left the slab migrate to the
 * owning neighbour (in place: leavers leave holes that arrivals and tail agents
 * fill).  Transport, default: PEER MEMORY -- the receive buffers are exchanged as
 * CUDA IPC handles at init and the push kernels (k_push_planes, k_push) store
 * straight into the neighbours' buffers over NVLink, publishing a sequence number
 * with st.release.sys and waiting (bounded) for the neighbours'; NCCL is then used
 * once, to exchange the handles.  Fallback (IPC unavailable, or option
 * "slab_p2p" = 0): NCCL grouped send/recv of the fixed-capacity buffers on the
 * simulation stream.  Buffers hold halo_cap messages / mig_cap agents per
 * direction (0 = automatic), true counts in a device-side header.
 * `nccl_unique_id` is the 128-byte ncclUniqueId made by rank 0
 * (fgpu_nccl_unique_id) and broadcast by the host program. */
int fgpu_nccl_unique_id(void *out128);
int fgpu_sim_slab_init(fgpu_sim *sim, int rank, int world, const void *nccl_unique_id, int halo_cap, int mig_cap);
/* {enabled, axis, nplanes, first owned plane, one past last owned plane, halo_cap, mig_cap, cells per plane} */
int fgpu_sim_slab_info(fgpu_sim *sim, int *out8);

#ifdef __cplusplus
}
#endif
#endif /* FGPU_H */
