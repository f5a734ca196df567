/*
 * numga_b200 — C ABI of the B200 (sm_100a) backend for numga's data-parallel hot path.
 *
 * The reference (EelcoHoogendoorn/numga) is pure Python and has no FFI; its plug-in boundary is
 * the duck-typed backend triple Context / MultiVector / Operator (numga/context.py:55-69,
 * numga/operator/abstract.py:12-60).  The entry points below are what a `numga/backend/b200/`
 * directory binds with ctypes: each one names the reference interface it
 * stands in for.
 *
 * Conventions
 *  - plain pointers and sizes only; no torch / C++ types cross this boundary;
 *  - every function returns NGB200_OK (0) or a negative error code and never throws;
 *    `ngb200_last_error()` gives the message of the last failure on the calling thread;
 *  - the library never allocates the caller's arrays and never synchronises the caller's stream
 *    (`*_launch`); `*_run_host` is the synchronous host-buffer variant;
 *  - device pointers are CUDA device pointers in the CURRENT context (the primary context torch
 *    or the CUDA runtime made current); `stream` is a CUstream / cudaStream_t, 0 = default stream;
 *  - strides are in ELEMENTS of the program dtype.  Multivector batches are stored
 *    structure-of-arrays: component c of item i lives at ptr[c * comp_stride + i * batch_stride]
 *    (batch_stride 1 for SoA; numga's default numpy order='F' is exactly this layout,
 *    numga/backend/numpy/context.py:12,21-26).  A broadcast operand has batch_stride 0.
 *  - there is NO CPU fallback: without a CUDA device the launch functions fail with
 *    NGB200_ERR_CUDA.
 */
#ifndef NUMGA_B200_H
#define NUMGA_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NGB200_OK 0
#define NGB200_ERR_INVALID (-1) /* malformed description / arguments          */
#define NGB200_ERR_COMPILE (-2) /* NVRTC rejected the generated kernel        */
#define NGB200_ERR_CUDA (-3)    /* driver error, or no CUDA device available  */
#define NGB200_ERR_IO (-4)      /* kernel cache directory not usable          */

/* element type of every operand of a program ("dtype" of the numga context) */
#define NGB200_F32 0
#define NGB200_F64 1

/* program flags */
#define NGB200_FAST 0u
/* Bit-faithful arithmetic: every multiply and add separately rounded, in program order, no FMA
 * contraction, IEEE division and square root.  Reproduces numpy's einsum result for unary and
 * binary operators bit for bit (hard part 1). */
#define NGB200_FAITHFUL 1u

/* operand addressing modes (fixed when the program is created) */
#define NGB200_VARYING 0   /* one multivector per batch item            */
#define NGB200_BROADCAST 1 /* one multivector shared by the whole batch */
#define NGB200_INDEXED 2   /* item i reads / writes row index[i] (gather / scatter) */
#define NGB200_TILED 4     /* inputs only: a PARTIALLY broadcast operand (numpy broadcasting between e.g. [c] and [2, c]
                              batches): item i reads row (i % inner_size) * batch_stride + (i / inner_size) *
                              outer_stride, so nothing has to be expanded in memory */
#define NGB200_REDUCED 3   /* outputs only: ONE multivector = the sum of the per-item results over the batch
                              (warp shuffle + shared + one atomicAdd per CTA and component); the caller zeroes it.
                              Used for the gradient of a broadcast operand and for sums over the batch axis.  The CTA
                              partial sums meet in floating-point atomicAdd, so the result is reproducible only up to
                              the rounding of a sum taken in a run-dependent order (relative ~1e-7 fp32 / 1e-16 fp64
                              per addend); not supported by ngb200_program_run_host. */

/* ---- straight-line program IR --------------------------------------------------------------
 * A program maps, independently for every batch item, the components of its input multivectors
 * to the components of its output multivectors.  Instruction k defines value k (SSA);
 * operands a, b, c refer to earlier values. */
enum ngb200_opcode {
    NGB200_OP_INPUT = 0,  /* a = input operand, b = component                    */
    NGB200_OP_CONST = 1,  /* a = index into consts[]                             */
    NGB200_OP_PARAM = 2,  /* a = index into the runtime scalar params of launch  */
    NGB200_OP_ADD = 3,
    NGB200_OP_SUB = 4,
    NGB200_OP_MUL = 5,
    NGB200_OP_DIV = 6,
    NGB200_OP_NEG = 7,
    NGB200_OP_SQRT = 8,
    NGB200_OP_RSQRT = 9,  /* 1 / sqrt(a), both IEEE-rounded                      */
    NGB200_OP_RCP = 10,   /* 1 / a                                               */
    NGB200_OP_ABS = 11,
    NGB200_OP_NAN_TO_NUM = 12, /* nan -> value b (a value id); +-inf -> +-max finite (numpy semantics) */
    NGB200_OP_FMA = 13,   /* a * b + c, single rounding (FAST programs only)     */
    NGB200_OP_EXP = 14,
    NGB200_OP_LOG = 15,
    NGB200_OP_SIN = 16,
    NGB200_OP_COS = 17,
    NGB200_OP_MIN = 18,
    NGB200_OP_MAX = 19,
    NGB200_OP_SELECT_GT0 = 20, /* a > 0 ? b : c                                   */
    NGB200_OP_ATAN2 = 21, /* atan2(a, b): angle of the point (b, a)               */
    NGB200_OP_PARTNER = 22, /* value a as held by the thread of item i ^ 1 (warp shuffle): PIPELINE PHASES ONLY, items come in
                               pairs (2k, 2k + 1), e.g. the two sides of a constraint (numga/examples/physics/core.py arrays [2, c]) */
    NGB200_OP_COUNT_
};

typedef struct ngb200_instr {
    int32_t op, a, b, c;
} ngb200_instr;

typedef struct ngb200_store {
    int32_t operand, component, value;
} ngb200_store;

typedef struct ngb200_program_desc {
    int32_t dtype; /* NGB200_F32 | NGB200_F64 */
    uint32_t flags;
    int32_t n_inputs;
    const int32_t *input_width; /* components per input multivector  */
    const int32_t *input_mode;  /* NGB200_VARYING | BROADCAST | INDEXED | TILED */
    int32_t n_outputs;
    const int32_t *output_width;
    const int32_t *output_mode; /* NGB200_VARYING | INDEXED | REDUCED */
    int32_t n_params;           /* runtime scalars (passed as double, cast to dtype) */
    int32_t n_consts;
    const double *consts;
    int32_t n_instr;
    const ngb200_instr *instr;
    int32_t n_stores;
    const ngb200_store *stores; /* output components not stored are written as 0 */
    int32_t vec;                /* batch items per thread: 0 = automatic, 1, 2 or 4 */
} ngb200_program_desc;

typedef struct ngb200_operand {
    void *ptr;
    int64_t comp_stride;
    int64_t batch_stride;
    const int64_t *index; /* NGB200_INDEXED operands only, else NULL */
    int64_t inner_size;   /* NGB200_TILED operands only, else 0 */
    int64_t outer_stride; /* NGB200_TILED operands only, else 0 */
} ngb200_operand;

typedef struct ngb200_program ngb200_program;

typedef struct ngb200_program_info {
    int32_t n_instr;      /* live instructions after dead-code elimination */
    int32_t n_uniform;    /* of which evaluated once per thread (broadcast-only) */
    int32_t vec;          /* items per thread of the contiguous kernel (4/2: 128/64-bit accesses; 1: plain accesses) */
    int32_t regs_vec;     /* registers per thread (0 until loaded on a device) */
    int32_t regs_scalar;
    int32_t cache_hit;    /* 1 if the cubin came from the on-disk cache */
    int64_t cubin_bytes;
} ngb200_program_info;

/* Library identification: "numga_b200 <version> sm_100a nvrtc <major>.<minor>". */
const char *ngb200_version(void);
const char *ngb200_last_error(void);

/* Directory of the persistent cubin cache (default: $NGB200_CACHE_DIR, else <library dir>/_kcache). */
int ngb200_set_cache_dir(const char *path);

/* Create the kernel for one symbolic operator from its integer term table.
 * Replaces: AbstractConcreteOperator.__init__ + precompute_sparse (numga/operator/abstract.py:21-24,
 * 119-151) and the per-backend operator classes (numga/backend/numpy/operator.py:102-160).
 * terms: n_terms rows of (arity + 2) int32: i_0 .. i_{arity-1}, o, coeff, in precompute_sparse order.
 * bcast_mask bit k set: input k is one multivector shared by the batch (numpy stride-0 broadcasting,
 * numga/backend/numpy/operator.py:40-49).  Compiles for sm_100a (NVRTC); needs no GPU. */
int ngb200_operator_create(const int32_t *terms, int32_t n_terms, int32_t arity, const int32_t *n_in,
                           int32_t n_out, int32_t dtype, uint32_t bcast_mask, uint32_t flags,
                           ngb200_program **out);

/* Create a fused kernel from a straight-line program (operator chains such as sandwich, normalized,
 * exp, motor_log; numga/multivector/extension/{logexp,norms,roots,inverse}.py run as ONE pass). */
int ngb200_program_create(const ngb200_program_desc *desc, ngb200_program **out);

void ngb200_program_destroy(ngb200_program *prog);
int ngb200_program_get_info(const ngb200_program *prog, ngb200_program_info *info);
/* Generated CUDA C++ source (for inspection / profiling with -lineinfo). */
const char *ngb200_program_source(const ngb200_program *prog);

/* Apply the program to `batch` items on the device.  Replaces XOperator.__call__
 * (numga/backend/numpy/operator.py:117-132).  Asynchronous on `stream`. */
int ngb200_program_launch(ngb200_program *prog, const ngb200_operand *inputs, const ngb200_operand *outputs,
                          const double *params, int64_t batch, void *stream);

/* Same, with HOST buffers: copies inputs to the device, runs, copies outputs back, pipelined in
 * chunks over several streams; returns when the outputs are complete.  Host buffers should be
 * page-locked (ngb200_host_alloc) for full PCIe bandwidth.  `device` is the CUDA ordinal;
 * chunk_items 0 = automatic. */
int ngb200_program_run_host(ngb200_program *prog, const ngb200_operand *inputs, const ngb200_operand *outputs,
                            const double *params, int64_t batch, int32_t device, int64_t chunk_items);

/* Raw-copy control for benchmarks: the chunked multi-stream H2D of in_bytes + D2H of out_bytes that
 * ngb200_program_run_host performs, with no kernel (chunk_bytes / streams 0 = the pipeline's defaults).
 * host_out receives unspecified data. */
int ngb200_host_copy_control(void *host_in, int64_t in_bytes, void *host_out, int64_t out_bytes, int64_t chunk_bytes,
                             int32_t streams, int32_t device);

/* Launch plans: the steady state of an eager backend (same program, same operand layout, fresh pointers every
 * call; XOperator.__call__ allocates its result each time, numga/backend/numpy/operator.py:40-49,117-132).
 * A plan freezes kernel choice, grid sizes and the parameter block for one layout + batch on the CURRENT device;
 * ngb200_plan_launch patches pointers and issues the kernel(s): no validation loops, no allocation.
 * Layouts are given as operands whose `ptr` is ignored; programs with INDEXED / TILED / REDUCED operands are not
 * plannable (NGB200_ERR_INVALID).  A plan must not outlive its program. */
typedef struct ngb200_plan ngb200_plan;
int ngb200_plan_create(ngb200_program *prog, const ngb200_operand *input_layout, const ngb200_operand *output_layout,
                       int64_t batch, ngb200_plan **out);
int ngb200_plan_launch(ngb200_plan *plan, void *const *input_ptrs, void *const *output_ptrs, const double *params,
                       void *stream);
void ngb200_plan_destroy(ngb200_plan *plan);

/* ---- pipelines: several programs fused into ONE kernel with on-chip exchange ----------------------------------
 * Replaces a CHAIN of operator applications with neighbour coupling that the reference runs as hundreds of array
 * passes (numga/examples/physics/core.py:80-88 `Body.integrate`: pre_integrate -> Constraint.apply per colour set
 * -> post_integrate -> Constraint.v_apply per colour set; 174 launches per sub-step).
 * Items are grouped into TILES: tile t holds items [t * items_per_tile[d], (t + 1) * items_per_tile[d]) of every
 * item domain d (e.g. bodies, constraints of colour set 0, constraints of colour set 1) and is processed by one CTA.
 * A phase is a straight-line program over the items of one domain; its operands are bound to GLOBAL arrays (HBM;
 * item i of the domain, or row index[i] through an int64 index operand = gather / scatter) or to SHARED buffers
 * (on-chip for the life of the tile, component-major; item i, or row index[i] - tile * items_per_tile of the
 * buffer's domain, which must fall inside the tile).  Phases are separated by CTA barriers; phases
 * [loop_begin, loop_end) repeat `repeat` times per launch.  Scatter destinations must be unique per phase. */
#define NGB200_BIND_GLOBAL 0
#define NGB200_BIND_SHARED 1

typedef struct ngb200_bind {
    int32_t kind;  /* NGB200_BIND_GLOBAL | NGB200_BIND_SHARED */
    int32_t slot;  /* global operand number | shared buffer number */
    int32_t index; /* -1: addressed by the item number; else the global operand number of an int64 index array over
                      the phase's domain */
} ngb200_bind;

typedef struct ngb200_phase_desc {
    ngb200_program_desc program; /* input_mode / output_mode / vec are ignored: addressing comes from the bindings */
    int32_t domain;
    const ngb200_bind *inputs;   /* program.n_inputs entries  */
    const ngb200_bind *outputs;  /* program.n_outputs entries */
} ngb200_phase_desc;

typedef struct ngb200_pipeline_desc {
    int32_t dtype;
    uint32_t flags;
    int32_t n_domains;
    const int32_t *items_per_tile;  /* per domain, 1..1024 */
    int32_t n_globals;
    const int32_t *global_width;    /* components (ignored for index arrays) */
    const int32_t *global_domain;   /* item domain of the array; -1 = one multivector shared by all items (read-only) */
    const int32_t *global_is_index; /* 1: int64 index array (ngb200_operand.ptr points at int64 data) */
    int32_t n_shared;
    const int32_t *shared_width;
    const int32_t *shared_domain;
    int32_t n_phases;
    const ngb200_phase_desc *phases;
    int32_t loop_begin, loop_end;
    int32_t n_params;               /* runtime scalars shared by all phases */
    int32_t threads_per_tile;       /* 0 = automatic */
} ngb200_pipeline_desc;

typedef struct ngb200_pipeline ngb200_pipeline;

typedef struct ngb200_pipeline_info {
    int32_t n_instr;          /* live instructions, all phases */
    int32_t threads_per_tile;
    int32_t min_blocks;       /* __launch_bounds__ minimum CTAs per SM */
    int32_t regs;             /* registers per thread (0 until loaded on a device) */
    int32_t grid;             /* persistent grid size (0 until loaded) */
    int32_t cache_hit;
    int64_t shared_bytes;     /* dynamic shared memory per CTA */
    int64_t spill_bytes;      /* ptxas spill stores (0 when the cubin came from the cache) */
    int64_t local_bytes;      /* local memory per thread reported by the driver */
} ngb200_pipeline_info;

int ngb200_pipeline_create(const ngb200_pipeline_desc *desc, ngb200_pipeline **out);
void ngb200_pipeline_destroy(ngb200_pipeline *pipe);
int ngb200_pipeline_get_info(const ngb200_pipeline *pipe, ngb200_pipeline_info *info);
const char *ngb200_pipeline_source(const ngb200_pipeline *pipe);
/* globals: one operand per global array (ptr, comp_stride; batch_stride must be 1; index arrays: ptr only).
 * domain_items: total item count per domain; n_tiles CTAs' worth of tiles are processed. Asynchronous on `stream`. */
int ngb200_pipeline_launch(ngb200_pipeline *pipe, const ngb200_operand *globals, const int64_t *domain_items,
                           int64_t n_tiles, int32_t repeat, const double *params, void *stream);

int ngb200_host_alloc(void **ptr, int64_t bytes);
int ngb200_host_free(void *ptr);

/* Synthetic data for benchmarks: fill n elements with i.i.d. normal(0, scale) from a counter-based
 * generator (same values for the same seed/offset on any GPU). */
int ngb200_fill_normal(void *ptr, int64_t n, uint64_t seed, uint64_t offset, double scale, int32_t dtype,
                       void *stream);

/* Number of kernels this library has launched in this process (bench.py reports it). */
int64_t ngb200_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* NUMGA_B200_H */
