/*
 * fgpu_model.h -- the model-plugin ABI.
 *
 * A FLAME GPU 1.5 model (XMLModelFile.xml + functions.c) is compiled by
 * `python -m ev_diffusion_b200.build` into one shared object which exports
 * `fgpu_model_get()`.  The descriptor it returns replaces what the reference
 * bakes into its XSLT-generated simulation.cu / FLAMEGPU_kernals.cu
 * (FLAMEGPU/templates/simulation.xslt:225-330 globals, :1402-1974 wrappers;
 * FLAMEGPU/templates/FLAMEGPU_kernals.xslt:1316-1387 GPUFLAME_<f> kernels).
 *
 * Plain C: pointers, sizes, function pointers.  No CUDA or C++ types.
 */
#ifndef FGPU_MODEL_H
#define FGPU_MODEL_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FGPU_MODEL_ABI_VERSION 12

/* One agent or message variable.  Agent lists keep the reference's SoA struct
 * layout (header.xslt:187-194): int _position[MAX]; int _scan_input[MAX];
 * T var[MAX]; ... so `list_offset` is offsetof(xmachine_memory_<A>_list, var). */
typedef struct fgpu_var {
    const char *name;
    const char *ctype;       /* "int", "float", "double", ... */
    int size;                /* bytes: 4 or 8 */
    int is_float;            /* 1 for float/double */
    size_t list_offset;      /* agents: byte offset of the column in the list struct;
                                messages: byte offset inside the packed record */
    double default_value;    /* XMML defaultValue or 0 (io.xslt:276-291) */
    int length;              /* 1 for a scalar; arrayLength of an agent array variable (header.xslt:187-194:
                                T var[MAX * length], element e of agent i at var[e * MAX + i]) */
} fgpu_var;

typedef struct fgpu_agent {
    const char *name;
    int buffer_size;         /* xmachine_memory_<A>_MAX */
    int nvars;
    const fgpu_var *vars;
    int nstates;
    const char *const *states;
    int initial_state;
    size_t list_bytes;       /* sizeof(xmachine_memory_<A>_list) */
    size_t position_offset;  /* offsetof(_position)  (= 0) */
    size_t scan_offset;      /* offsetof(_scan_input) */
    int carry_quads;         /* sizeof(packed record of all scalar variables) / 16: the record a message producer writes
                                beside its message so that the build can deliver the agents' variables in cell order
                                (fgpu_launch.carry_out / carry_in); 0 = the plugin does not support it */
} fgpu_agent;

enum { FGPU_PART_NONE = 0, FGPU_PART_SPATIAL = 1 };

typedef struct fgpu_message {
    const char *name;
    int buffer_size;         /* xmachine_message_<M>_MAX */
    int nvars;
    const fgpu_var *vars;
    int rec_quads;           /* record = sizeof(xmachine_message_<M>) / 16; lists are arrays of records */
    int partitioning;
    float radius;
    float min_bounds[3];
    float max_bounds[3];
    int dims[3];             /* (int)round((max-min)/radius) in float, simulation.xslt:540-542 */
    int grid_size;           /* xmachine_message_<M>_grid_size, header.xslt:102-105 */
    int is_2d;               /* FLAMEGPU_kernals.xslt:1052 */
    int radix_bits;          /* simulation.xslt:282 */
    int xvar, yvar, zvar;    /* indices of the position variables */
} fgpu_message;

/* Arguments of one agent-function launch.  Everything the reference passes
 * through kernel arguments, __constant__ symbols and bound textures
 * (simulation.xslt:1597-1691) travels in this one struct, by value. */
typedef struct fgpu_launch {
    void *agents;                 /* working list d_<A>s */
    int n;                        /* working count (exact, host-known like h_xmachine_memory_<A>_count) */
    const int *d_n;               /* slab mode: exact count on the device, `n` is then an upper bound; else NULL */
    const unsigned int *order;    /* processing order: thread t runs list index order[t]; NULL = identity.
                                     Always a permutation of [0,n): results never depend on it. */
    void *new_agents;             /* d_<B>s_new (sparse newborn list) or NULL */
    /* input message (sorted, cell-indexed) */
    const void *in_recs;          /* xmachine_message_<M>[count], cell-sorted */
    const unsigned int *in_cell_start;  /* CSR: grid_size+1 entries */
    int in_count;                 /* h_message_<M>_count */
    const int *d_in_count;        /* slab mode: exact message count on the device, or NULL */
    /* output message (unsorted staging + cell keys) */
    void *out_recs;
    unsigned int *out_keys;
    int *out_flags;               /* optional_message output: add_<M>_message sets out_flags[slot] = 1 (zeroed before
                                     the launch); NULL for single_message */
    int out_base;                 /* d_message_<M>_count before this function (krn:385) */
    /* rand48 (simulation.xslt:661-688) */
    void *seeds;                  /* uint2[buffer_size_MAX], .x low 24 bits, .y high 24 bits */
    unsigned int Ax, Ay, Cx, Cy;
    /* carried agent records (16 * carry_quads bytes each: every scalar variable, XML order, natural alignment) */
    const void *carry_in;         /* consumer: record t = the variables of list index order[t], i.e. in cell order, as the
                                     producer of the input message left them; NULL = read the SoA list at order[t] */
    void *carry_out;              /* producer: the wrapper stores record `index` after the script returns; NULL = no */
    /* launch window over the cell-order positions (functions with a spatial input message, order != NULL): slab mode
     * runs the agents of the interior planes while the halo planes are still in flight, the boundary planes after.
     *   w0 = in_cell_start[win_cell_lo] - win_base,  w1 = in_cell_start[win_cell_hi] - win_base  (read on the device)
     *   win_mode 0: every position;  1: positions [w0, w1);  2: positions [0, w0) and [w1, n) */
    int win_mode;
    unsigned int win_cell_lo, win_cell_hi, win_base;
    int launch_n;                 /* win_mode != 0: number of threads to launch (a host-side upper bound) */
    int pdl;                      /* launch with the programmatic-stream-serialization attribute (the wrapper kernel starts
                                     with griddepcontrol.wait, so it is safe either way) */
    /* own-message shortcut (fgpu_function.mirror_producer): sorted message mirror_base + t is the message list index
     * order[t] emitted, and it carries some of the agent's variables unchanged -- the wrapper reads those from that one
     * record (coalesced) instead of gathering them from the SoA list at a random index */
    int mirror;
    unsigned int mirror_base;     /* record index of the local block inside in_recs (0 on one GPU; slab mode: behind the low halo) */
} fgpu_launch;

typedef void (*fgpu_launch_fn)(const fgpu_launch *args, void *cuda_stream);

/* Function-condition / global-condition filter (FLAMEGPU_kernals.xslt:162-210). */
typedef struct fgpu_filter {
    void *current;                /* d_<A>s_<currentState> */
    void *working;                /* d_<A>s (function condition only) */
    int n;
    int *flags_current;           /* -> current->_scan_input */
    int *flags_working;           /* -> working->_scan_input */
} fgpu_filter;
typedef void (*fgpu_filter_fn)(const fgpu_filter *args, void *cuda_stream);

enum { FGPU_COND_NONE = 0, FGPU_COND_FUNCTION = 1, FGPU_COND_GLOBAL = 2 };
enum { FGPU_OUT_SINGLE = 0, FGPU_OUT_OPTIONAL = 1 };

typedef struct fgpu_function {
    const char *name;
    int agent;
    int current_state, next_state;
    int input_message;            /* -1 if none */
    int output_message;           /* -1 if none */
    int output_type;
    int xagent_output;            /* agent index or -1 */
    int xagent_output_state;
    int reallocate;               /* gpu:reallocate == 'true' */
    int swappable;                /* the literal test of simulation.xslt:1945 */
    int rng;
    int condition;                /* FGPU_COND_* */
    int gc_max_iterations;
    int gc_must_evaluate_to;
    fgpu_launch_fn launch;
    fgpu_filter_fn filter;        /* NULL unless condition != NONE */
    int block_size;               /* threads per block the launcher uses */
    int writes_position;          /* the script may modify x, y or z (slab mode: agents may change owner) */
    int mirror_producer;          /* >= 0: index of the one function whose (spatial, single_message) output this function consumes
                                     AND whose add_<M>_message call passes variables of this agent type through unchanged: when
                                     that function ran last on the list this one runs on, the runtime may set fgpu_launch.mirror.
                                     -1: never */
} fgpu_function;

typedef struct fgpu_layer {
    int nfunctions;
    const int *functions;
} fgpu_layer;

typedef struct fgpu_constant {
    const char *name;
    const char *ctype;
    int size;                     /* bytes of one element */
    int length;                   /* array length (1 for scalars) */
    void (*set)(const void *host_value);    /* set_<name>(): cudaMemcpyToSymbol, hdr:771 */
    const void *(*get)(void);               /* get_<name>(): host shadow copy */
} fgpu_constant;

typedef struct fgpu_model {
    int abi_version;
    const char *name;
    int buffer_size_max;          /* buffer_size_MAX, header.xslt:71-75 */
    int nagents;
    const fgpu_agent *agents;
    int nmessages;
    const fgpu_message *messages;
    int nfunctions;
    const fgpu_function *functions;
    int nlayers;
    const fgpu_layer *layers;
    int nconstants;
    const fgpu_constant *constants;
    void (*init_constants)(void);  /* initEnvVars(): XMML defaultValues, io.xslt:201-209 */
    int ninit_functions;
    void (*const *init_functions)(void);
    int nstep_functions;
    void (*const *step_functions)(void);
    int nexit_functions;
    void (*const *exit_functions)(void);
    /* Makes `sim` the simulation the plugin's Face-2 shim (initialise, singleIteration,
     * get_agent_<A>_<S>_count ... header.xslt:531-645) and its init/step/exit functions act on. */
    void (*bind)(void *sim);
    /* Agent id generation (generate_<A>_id / set_initial_<A>_id, header.xslt:500-512).  The counters live in the
     * plugin (a __device__ symbol and its host shadow); the runtime calls these where the reference synchronises
     * them: after the initial states are read (io.xslt:599-616: first id = largest id read + 1), before and after
     * the step / init / exit functions (simulation.xslt:918-960).  NULL when no agent has an integer `id`. */
    const char *const *init_function_names;   /* for the "Instrumentation: <name> = ..." lines; may be NULL */
    const char *const *step_function_names;
    const char *const *exit_function_names;
    void (*ids_after_load)(void *sim);
    void (*ids_to_host)(void);
    void (*ids_to_device)(void);
} fgpu_model;

/* The single symbol a model plugin exports. */
const fgpu_model *fgpu_model_get(void);

#ifdef __cplusplus
}
#endif
#endif /* FGPU_MODEL_H */
