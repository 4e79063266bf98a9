/*
 * dxo.h — C ABI of the B200-native Draco attribute-encoding path.
 *
 * This header is the drop-in boundary for draco-oxide's
 *   pub fn encode<W: ByteWriter>(mesh: Mesh, writer: &mut W, cfg: Config) -> Result<(), Err>
 *   (reference: draco-oxide/src/encode/mod.rs:59-97)
 * and for the per-primitive call the glTF transcoder makes
 *   (reference: draco-oxide/src/io/gltf/encode.rs:932-955).
 *
 * Plain C types only: pointers, sizes, fixed-width integers. No CUDA, torch or
 * C++ types appear in any signature. All inputs are borrowed for the duration
 * of the call; outputs are owned by the library until dxo_free_bytes().
 *
 * The byte stream produced is the Draco v2.2 Edgebreaker mesh stream of the
 * reference, byte for byte (layout: DESIGN.md "Byte layout").
 *
 * There is no CPU fallback: every entry point that does attribute work fails
 * with DXO_ERR_NO_DEVICE / DXO_ERR_CUDA when no sm_100a device is usable.
 */
#ifndef DXO_H_
#define DXO_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DXO_ABI_VERSION 1

/* ---- enums (numeric ids are the ones the reference writes to the stream) ---- */

/* AttributeType::get_id — draco-oxide/src/core/attribute/mod.rs:652-665 */
enum dxo_attribute_type {
  DXO_ATT_POSITION = 0,
  DXO_ATT_NORMAL = 1,
  DXO_ATT_COLOR = 2,
  DXO_ATT_TEXCOORD = 3,
  DXO_ATT_CUSTOM = 4,
  DXO_ATT_TANGENT = 5,
  DXO_ATT_MATERIAL = 6,
  DXO_ATT_JOINT = 7,
  DXO_ATT_WEIGHT = 8
};

/* ComponentDataType::get_id — draco-oxide/src/core/attribute/mod.rs:565-579 */
enum dxo_component_type {
  DXO_U8 = 1,
  DXO_I8 = 2,
  DXO_U16 = 3,
  DXO_I16 = 4,
  DXO_U32 = 5,
  DXO_I32 = 6,
  DXO_U64 = 7,
  DXO_I64 = 8,
  DXO_F32 = 9,
  DXO_F64 = 10
};

/* AttributeDomain — draco-oxide/src/core/attribute/mod.rs:700-716 */
enum dxo_domain { DXO_DOMAIN_POSITION = 0, DXO_DOMAIN_CORNER = 1 };

/* ---- error codes: one per leaf of encode::Err (encode/mod.rs:44-56) ---- */
enum dxo_status {
  DXO_OK = 0,
  DXO_ERR_INVALID_ARGUMENT = -1,     /* NULL pointer, bad sizes, index out of range */
  DXO_ERR_UNSUPPORTED_INPUT = -2,    /* inputs on which the reference panics / unimplemented!() */
  DXO_ERR_UNSUPPORTED_DATA_TYPE = -3,/* attribute_encoder.rs:33  Err::UnsupportedDataType */
  DXO_ERR_UNSUPPORTED_NUM_COMPONENTS = -4, /* attribute_encoder.rs:35 */
  DXO_ERR_TOO_MANY_ATTRIBUTES = -5,  /* connectivity/mod.rs:110 TooManyConnectivityAttributes */
  DXO_ERR_RANS_INVALID_SYMBOL = -6,  /* entropy/rans.rs:261 InvalidSymbolIndex */
  DXO_ERR_RANS_STATE_TOO_LARGE = -7, /* entropy/rans.rs:265 StateTooLarge */
  DXO_ERR_RANS_FREQ_TABLE = -8,      /* shared/entropy/mod.rs:69 FrequencyCountNotCompatibleWithRansPrecision */
  DXO_ERR_ZERO_NORMAL = -9,          /* geom.rs:45 assert!(v != zero) */
  DXO_ERR_UNUSED_VERTICES = -10,     /* corner_table/mod.rs:105-108 panic */
  DXO_ERR_NO_DEVICE = -20,           /* no usable CUDA device: there is no CPU fallback */
  DXO_ERR_CUDA = -21,                /* a CUDA runtime call or kernel failed */
  DXO_ERR_OUT_OF_MEMORY = -22,
  DXO_ERR_INTERNAL = -99
};

/* ---- mesh handed over the boundary ----
 * Mirrors Mesh{faces, attributes} and Attribute{id, buffer, att_type, domain,
 * parents, point_to_att_val_map} (core/mesh/mod.rs:13-23, core/attribute/mod.rs:26-49).
 * A Rust shim converts usize indices to u32 and fills these from &Mesh.
 */
typedef struct dxo_attribute {
  uint32_t att_type;        /* enum dxo_attribute_type */
  uint32_t component_type;  /* enum dxo_component_type */
  uint32_t num_components;  /* 1..4 */
  uint32_t domain;          /* enum dxo_domain */
  uint32_t unique_id;       /* AttributeId */
  uint32_t num_parents;
  const uint32_t* parent_ids;       /* AttributeId of each parent */
  uint64_t num_unique_values;       /* Attribute::num_unique_values() */
  const void* values;               /* AoS little-endian, num_unique_values * num_components components */
  uint64_t num_points;              /* Attribute::len() */
  const uint32_t* point_to_value;   /* point -> unique value index; NULL = identity map */
} dxo_attribute;

typedef struct dxo_mesh {
  uint64_t num_faces;
  const uint32_t* faces;            /* 3 point indices per face */
  uint32_t num_attributes;
  const dxo_attribute* attributes;  /* attributes[0] is the position attribute (MeshBuilder::get_sorted_attributes) */
} dxo_mesh;

/* encode::Config (encode/mod.rs:22-43) has only default(); the quantization bit
 * counts are constants in the reference (portabilization/mod.rs:116-142). They
 * are parameters here; only 11/10 is reference behaviour. Octahedral normal
 * quantization is fixed at 8 bits (the reference hard-codes 255/127). */
/* Resident sessions replay the device step as one CUDA graph from their second run on (one driver call per step instead
 * of ~80). Off by default: the flag copies for the host-coded side streams leave later from inside a graph. */
#define DXO_FLAG_GRAPH_REPLAY 1u

typedef struct dxo_config {
  uint32_t abi_version;       /* DXO_ABI_VERSION */
  uint32_t position_bits;     /* default 11 */
  uint32_t texcoord_bits;     /* default 10 */
  uint32_t generic_bits;      /* default 11: other quantized attribute types */
  int32_t  device;            /* CUDA device ordinal; -1 = current device */
  uint32_t flags;             /* DXO_FLAG_* bits, 0 by default */
} dxo_config;

typedef struct dxo_bytes {
  uint8_t* data;
  size_t len;
} dxo_bytes;

/* Fills *cfg with the reference defaults (encode::Config::default()). */
void dxo_config_default(dxo_config* cfg);

/* encode::encode — one mesh, host buffers in, owned byte buffer out. */
int dxo_encode(const dxo_mesh* mesh, const dxo_config* cfg, dxo_bytes* out);

/* The transcoder loop (io/gltf/encode.rs:941-953 called once per primitive):
 * n independent meshes, sharded over GPUs [first_gpu, first_gpu+num_gpus) with
 * no exchange between them. outs[i] corresponds to meshes[i]; statuses[i]
 * (optional) receives the per-mesh status. Returns the first non-OK status. */
int dxo_encode_batch(const dxo_mesh* meshes, size_t n, const dxo_config* cfg,
                     dxo_bytes* outs, int* statuses, int first_gpu, int num_gpus);

void dxo_free_bytes(dxo_bytes* b);
const char* dxo_strerror(int status);

/* Number of visible CUDA devices (0 when none / no driver). */
int dxo_device_count(void);

/* ---- resident sessions: the measured hot path with inputs already in HBM ----
 * A session owns the host-side connectivity results (corner tables, Edgebreaker
 * bytes, attribute sequences) and the device copies of every array the
 * attribute kernels read. dxo_session_run() executes only the device hot path
 * (quantize -> predict -> symbolize -> histogram -> table -> rANS) plus the
 * D2H of its results, and reassembles the stream. The mesh and the arrays it
 * points to are borrowed for the life of the session: they must stay valid and
 * unchanged until dxo_session_destroy (without a position map the faces ARE the
 * corner -> vertex table the session's traces read). */
typedef struct dxo_session dxo_session;

int dxo_session_create(const dxo_mesh* mesh, const dxo_config* cfg, dxo_session** out);
/* Host-only part of encode(): header + Edgebreaker connectivity + attribute-section
 * headers ("head_bytes") and the per-attribute traversal sequences, readable through
 * dxo_session_trace_get. Needs no device; dxo_session_run on it returns DXO_ERR_NO_DEVICE. */
int dxo_connectivity_create(const dxo_mesh* mesh, const dxo_config* cfg, dxo_session** out);
/* Runs the device hot path once. out may be NULL (results are discarded after
 * the D2H completes). */
int dxo_session_run(dxo_session* s, dxo_bytes* out);
/* Runs the device hot path `steps` times back to back (bench loop). *ms_total is the
 * CUDA-event time, on the launching stream, from before the first launch of the first
 * step to after the last step's results reached the host. */
int dxo_session_run_steps(dxo_session* s, uint32_t steps, float* ms_total, uint64_t* launches_total);
void dxo_session_destroy(dxo_session* s);

/* Timing of the last dxo_session_run / dxo_encode on this thread, measured with
 * CUDA events on the stream the kernels were launched on. */
typedef struct dxo_kernel_time {
  const char* name;        /* kernel label, static storage */
  float ms;                /* device time of this launch */
  uint64_t algorithmic_bytes; /* bytes per DESIGN.md "Algorithmic bytes" for this launch */
} dxo_kernel_time;

typedef struct dxo_timing {
  float device_ms;         /* first kernel start -> last kernel end */
  float host_connectivity_ms; /* corner tables + Edgebreaker + sequencer (host) */
  float h2d_ms;
  float d2h_ms;
  float total_ms;          /* wall clock of the call */
  uint64_t h2d_bytes;
  uint64_t d2h_bytes;
  uint32_t num_launches;   /* kernels launched by this library in the call */
  uint32_t num_kernels;    /* entries valid in kernels[] */
  dxo_kernel_time kernels[64];
} dxo_timing;

/* Per-kernel event timing: 0 = off, 1 = event records between launches (attributes still overlap on their
 * streams), 2 = as 1 with the attributes run one after the other, so every kernel is timed alone. */
void dxo_set_profiling(int mode);
int dxo_last_timing(dxo_timing* out);

/* ---- stage access for parity tests (GPU results, copied back) ----
 * After dxo_session_run with tracing enabled, intermediate device results can
 * be read by key ("att0.quantized", "att0.symbols", "att1.flips", ...). The
 * pointer stays valid until the next run or session destroy. */
void dxo_session_set_trace(dxo_session* s, int enabled);
int dxo_session_trace_get(dxo_session* s, const char* key, const void** data, uint64_t* nbytes);

/* encode_symbols(symbols, _, SymbolEncodingMethod::DirectCoded, writer)
 * (encode/entropy/symbol_coding.rs:17-55) on the device: histogram, probability table,
 * rANS. *out receives exactly the bytes the reference writes. kernel_ms (optional, 3
 * floats) receives the device time of the histogram, table and rANS kernels. */
int dxo_encode_symbols(const uint32_t* symbols, uint64_t n, int device, dxo_bytes* out, float* kernel_ms);

/* RabsCoder::new / write / flush (encode/entropy/rans.rs:71-127) over a whole bit sequence, in the order given: the host
 * coder behind every binary side stream (seam flags, flips, orientations). bits: n bytes, zero / non-zero. zero_prob in
 * [1, 255]. mode 0 = the bit-by-bit coder, 1 = the coder the encoder uses (table-driven for long, heavily skewed
 * streams; same bytes). Host only: needs no device. */
int dxo_encode_bits(const uint8_t* bits, uint64_t n, uint8_t zero_prob, int mode, dxo_bytes* out);

/* Corner-table build on the device (half-edge matching by radix sort;
 * replaces CornerTable::compute_table, core/corner_table/mod.rs:252-340).
 * vertex_of_corner: 3*num_faces vertex ids. opposite_out: 3*num_faces entries,
 * 0xFFFFFFFF = none. *exact_out = 1 when the mesh is on the manifold fast path
 * (result equals the reference's order-dependent matcher); 0 means the caller
 * must use the sequential matcher. */
int dxo_corner_table_opposites(const uint32_t* vertex_of_corner, uint64_t num_faces,
                               uint32_t* opposite_out, int* exact_out, int device);

/* ---- mesh build: the step before the boundary ----
 * dxo_mesh_build replaces MeshBuilder::add_attribute + MeshBuilder::build (core/mesh/builder.rs:30-125: value dedup
 * per attribute = Attribute::from / remove_duplicate_values, core/attribute/mod.rs:87-103, 394-452; position attribute
 * first; merge of points whose values agree in every attribute; removal of degenerate faces and unused points).
 * per_point_attributes[i]: `values` holds `num_unique_values` per-POINT values (one per point, all attributes the same
 * count); point_to_value and unique_id are ignored (ids are assigned in insertion order, parents refer to them).
 * The duplicate searches (all-pairs / hashing in the reference) run on the device as radix sorts with first-occurrence
 * numbering. Attributes of different lengths are reported as DXO_ERR_UNSUPPORTED_INPUT. */
typedef struct dxo_built_mesh dxo_built_mesh;
int dxo_mesh_build(const uint32_t* faces, uint64_t num_faces, const dxo_attribute* per_point_attributes,
                   uint32_t num_attributes, int device, dxo_built_mesh** out);
/* Borrowed view (valid until dxo_built_mesh_free), ready for dxo_encode. */
int dxo_built_mesh_view(const dxo_built_mesh* mesh, dxo_mesh* out);
void dxo_built_mesh_free(dxo_built_mesh* mesh);
/* Attribute::remove_duplicate_values alone: out_map[n] = point -> unique value (first-occurrence order),
 * out_first_index[u] = first point holding unique value u (u < *out_num_unique; the array needs n entries).
 * Equality is the typed `==`: -0.0 equals +0.0, a NaN equals nothing. */
int dxo_dedup_values(const void* values, uint64_t n, uint32_t component_type, uint32_t num_components, int device,
                     uint32_t* out_map, uint32_t* out_first_index, uint64_t* out_num_unique);

/* Accessor bounds of the glTF writer around encode(): compute_vec3_bounds / compute_vec4_bounds,
 * io/gltf/encode.rs:815-899 — per-component minimum and maximum over the values of all points (values[point_to_value[p]],
 * or values[p] when point_to_value is NULL and num_points is ignored), f32::min / f32::max semantics: NaNs are skipped, a
 * component that holds only NaNs yields NaN. num_components <= 4. With zero points the outputs are left untouched (the
 * reference returns empty vectors). A mix of -0.0 and +0.0 at a bound is unpinned in the reference (LLVM minnum); here
 * -0.0 < +0.0. Runs on the device (no CPU fallback). */
int dxo_attribute_bounds(const float* values, uint64_t num_values, uint32_t num_components, const uint32_t* point_to_value,
                         uint64_t num_points, int device, float* out_min, float* out_max);

/* ---- GLB assembly around the batch entry ----
 * What the reference's glTF writer does with every encoded primitive (io/gltf/encode.rs:932-1097 add_draco_mesh_internal,
 * :362-415 write_glb_format): the Draco streams go into the GLB BIN chunk (each padded to 4 bytes, one bufferView per
 * primitive), indices and POSITION / NORMAL / TEXCOORD_0 get placeholder accessors (POSITION with its true bounds), every
 * primitive carries KHR_draco_mesh_compression {bufferView, attributes}. streams[i] = the stream of meshes[i] (NULL: the
 * primitives are encoded here with dxo_encode_batch on GPUs [first_gpu, first_gpu + num_gpus)). The scene is the minimal
 * one (a mesh and a node per primitive); glTF input parsing, materials and textures stay with the caller. */
int dxo_encode_glb(const dxo_mesh* meshes, size_t n, const dxo_config* cfg, const dxo_bytes* streams, dxo_bytes* glb_out,
                   int first_gpu, int num_gpus);

#ifdef __cplusplus
}
#endif
#endif /* DXO_H_ */
