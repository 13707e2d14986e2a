// Whole-network entry points for hosts without Python (s2v_lnet_forward / s2v_dnet_forward /
// s2v_*_workspace_bytes).  The layer plan of a network at one batch size - the ordered list of C-ABI calls that
// models/LNet.py:122-139 / models/DNet.py:20-28 expand to, with folded + packed weights - is written once by the Python
// package (plan_export.py) into a relocatable file: every device pointer is stored as (arena, offset) with arena 0 = the
// constant image (weights, epilogue vectors, pointer tables) and arena 1 = the workspace (activations, statistics, I/O
// slots).  This file loads such a plan on the host, binds it to two caller-owned device buffers and replays it on a stream:
// no allocation on the device, no synchronisation, no Python.
//
// File layout (little endian):
//   "S2VPLAN1" | u32 version | u32 n_ops | u64 const_bytes | u64 workspace_bytes | u32 n_io | u32 n_init | u32 n_tab | u32 0
//   io   x n_io  : char name[32] | u64 offset | u64 bytes | u32 is_output | u32 0
//   init x n_init: u64 offset | u64 bytes | u32 kind (0 zero, 1 fill float32) | f32 value
//   tab  x n_tab : u64 offset (arena 0) | u32 bytes | u32 n_reloc | blob (padded to 8) | reloc x n_reloc
//   op   x n_ops : char fn[40] | u32 n_args | u32 0 | arg x n_args
//   arg          : u32 kind | u32 bytes | payload (padded to 8) [| u32 n_reloc | u32 0 | reloc x n_reloc   for kind BLOB]
//        kind 0 I64 (8 B) | 1 F64 (8 B) | 2 PTR (reloc: u32 at = 0, u32 arena, u64 offset) | 3 BLOB (struct passed by address)
//        | 4 NULL pointer
//   reloc        : u32 byte offset in the blob | u32 arena | u64 offset
//   const image  : const_bytes raw bytes
#include <cstdio>
#include <cstring>
#include <string>
#include <type_traits>
#include <utility>
#include <vector>

#include "common.cuh"

namespace s2v {
namespace plan {

enum ArgKind : uint32_t { kI64 = 0, kF64 = 1, kPtr = 2, kBlob = 3, kNull = 4 };
constexpr uint32_t kArenaConst = 0, kArenaWork = 1;

struct Reloc { uint32_t at, arena; uint64_t off; };
struct Arg {
  uint32_t kind = kNull;
  int64_t i = 0;
  double f = 0.0;
  void* ptr = nullptr;                 // kPtr after bind
  Reloc rel{};                         // kPtr
  std::vector<uint64_t> blob;          // kBlob (8-byte aligned storage)
  uint32_t blob_bytes = 0;
  std::vector<Reloc> relocs;           // kBlob
};
struct Io { char name[32]; uint64_t off, bytes; uint32_t is_output; };
struct Init { uint64_t off, bytes; uint32_t kind; float value; };
struct Tab { uint64_t off; uint32_t bytes; std::vector<uint64_t> blob; std::vector<Reloc> relocs; };
typedef int (*CallFn)(const void* fn, const Arg* a, void* stream);
struct Op { std::string name; const void* fn = nullptr; CallFn call = nullptr; uint32_t n_params = 0; std::vector<Arg> args; };

}  // namespace plan
}  // namespace s2v

struct s2v_plan {
  std::vector<s2v::plan::Op> ops;
  std::vector<s2v::plan::Io> io;
  std::vector<s2v::plan::Init> inits;
  std::vector<s2v::plan::Tab> tabs;
  std::vector<uint8_t> image;          // constant arena as stored in the file
  uint64_t const_bytes = 0, work_bytes = 0;
  uint8_t* cdev = nullptr;             // bound arenas (caller-owned)
  uint8_t* wdev = nullptr;
};

namespace s2v {
namespace plan {

// ---- generic unpacking: the parameter types of the C-ABI function decide how a stored argument is passed -------------
template <class T>
static T get_arg(const Arg& a) {
  if constexpr (std::is_pointer_v<T>) {
    using P = std::remove_cv_t<std::remove_pointer_t<T>>;
    if constexpr (std::is_same_v<P, s2v_view> || std::is_same_v<P, s2v_conv>)
      return a.kind == kBlob ? reinterpret_cast<T>(const_cast<uint64_t*>(a.blob.data())) : nullptr;   // host struct, by address
    else
      return a.kind == kPtr ? reinterpret_cast<T>(a.ptr) : nullptr;                                  // device pointer
  } else if constexpr (std::is_floating_point_v<T>) {
    return (T)(a.kind == kF64 ? a.f : (double)a.i);
  } else {
    return (T)(a.kind == kI64 ? a.i : (int64_t)a.f);
  }
}
template <size_t I, size_t N, class T>
static T pick(const Arg* a, void* stream) {
  if constexpr (I + 1 == N) return reinterpret_cast<T>(stream);       // the last parameter of every launcher is the stream
  else return get_arg<T>(a[I]);
}
template <class... A, size_t... I>
static int invoke(int (*fn)(A...), const Arg* a, void* stream, std::index_sequence<I...>) {
  return fn(pick<I, sizeof...(A), A>(a, stream)...);
}
template <class... A>
static int thunk(const void* fn, const Arg* a, void* stream) {
  return invoke(reinterpret_cast<int (*)(A...)>(const_cast<void*>(fn)), a, stream, std::index_sequence_for<A...>{});
}
struct Entry { const char* name; const void* fn; CallFn call; uint32_t n_params; };
template <class... A>
static Entry entry(const char* name, int (*fn)(A...)) {
  return Entry{name, reinterpret_cast<const void*>(fn), &thunk<A...>, (uint32_t)sizeof...(A) - 1};
}
#define S2V_PLAN_FN(f) entry(#f, &f)
static const Entry* find_fn(const char* name) {
  // every stream-ordered launcher a layer plan may contain (the ops of speech-to-video-mpp_b200/ops.py)
  static const Entry table[] = {
      S2V_PLAN_FN(s2v_conv_tc), S2V_PLAN_FN(s2v_conv_head), S2V_PLAN_FN(s2v_conv_simt), S2V_PLAN_FN(s2v_chan_stats),
      S2V_PLAN_FN(s2v_ln2d_finalize), S2V_PLAN_FN(s2v_ln2d_finalize_totals), S2V_PLAN_FN(s2v_adain_finalize),
      S2V_PLAN_FN(s2v_affine_act), S2V_PLAN_FN(s2v_affine_act2), S2V_PLAN_FN(s2v_adain_fused), S2V_PLAN_FN(s2v_token_layernorm),
      S2V_PLAN_FN(s2v_add), S2V_PLAN_FN(s2v_reflect_border), S2V_PLAN_FN(s2v_rfft2), S2V_PLAN_FN(s2v_irfft2),
      S2V_PLAN_FN(s2v_attention), S2V_PLAN_FN(s2v_pack_nchw_f32), S2V_PLAN_FN(s2v_unpack_to_nchw_f32),
      S2V_PLAN_FN(s2v_grouped_linear), S2V_PLAN_FN(s2v_resize_bilinear), S2V_PLAN_FN(s2v_resize_planes_f32),
      S2V_PLAN_FN(s2v_style_demod), S2V_PLAN_FN(s2v_style_epilogue), S2V_PLAN_FN(s2v_to_rgb),
      S2V_PLAN_FN(s2v_reflect_pad_nchw_f32), S2V_PLAN_FN(s2v_mean_over_w), S2V_PLAN_FN(s2v_flow_warp_f32),
      S2V_PLAN_FN(s2v_glue_fake_to_face_f32)};
  for (const Entry& e : table)
    if (strcmp(e.name, name) == 0) return &e;
  return nullptr;
}
#undef S2V_PLAN_FN

// ---- bounded reader --------------------------------------------------------------------------------------------------
struct Reader {
  const uint8_t* p; size_t n, at = 0; bool ok = true;
  bool take(void* dst, size_t k) {
    if (!ok || k > n - at) { ok = false; return false; }
    memcpy(dst, p + at, k); at += k; return true;
  }
  template <class T> T get() { T v{}; take(&v, sizeof(T)); return v; }
  bool skip(size_t k) { if (!ok || k > n - at) { ok = false; return false; } at += k; return true; }
};
static size_t pad8(size_t k) { return (k + 7) & ~(size_t)7; }

static bool read_relocs(Reader& r, uint32_t count, uint32_t blob_bytes, const s2v_plan& pl, std::vector<Reloc>& out) {
  if (count > (1u << 20)) return false;
  out.resize(count);
  for (Reloc& q : out) {
    q.at = r.get<uint32_t>(); q.arena = r.get<uint32_t>(); q.off = r.get<uint64_t>();
    if (!r.ok || q.arena > kArenaWork || (uint64_t)q.at + 8 > blob_bytes || (q.at & 7)) return false;
    if (q.off > (q.arena == kArenaConst ? pl.const_bytes : pl.work_bytes)) return false;
  }
  return true;
}

static int parse(const uint8_t* data, size_t size, s2v_plan& pl) {
  Reader r{data, size};
  char magic[8];
  if (!r.take(magic, 8) || memcmp(magic, "S2VPLAN1", 8) != 0) return S2V_EINVAL;
  const uint32_t version = r.get<uint32_t>(), n_ops = r.get<uint32_t>();
  pl.const_bytes = r.get<uint64_t>(); pl.work_bytes = r.get<uint64_t>();
  const uint32_t n_io = r.get<uint32_t>(), n_init = r.get<uint32_t>(), n_tab = r.get<uint32_t>();
  r.get<uint32_t>();
  if (!r.ok || version != 1 || n_ops > (1u << 20) || n_io > 64 || n_init > (1u << 16) || n_tab > (1u << 12)) return S2V_EINVAL;
  if (pl.const_bytes > size) return S2V_EINVAL;
  pl.io.resize(n_io);
  for (Io& e : pl.io) {
    r.take(e.name, 32); e.name[31] = 0;
    e.off = r.get<uint64_t>(); e.bytes = r.get<uint64_t>(); e.is_output = r.get<uint32_t>(); r.get<uint32_t>();
    if (!r.ok || e.off > pl.work_bytes || e.bytes > pl.work_bytes - e.off) return S2V_EINVAL;
  }
  pl.inits.resize(n_init);
  for (Init& e : pl.inits) {
    e.off = r.get<uint64_t>(); e.bytes = r.get<uint64_t>(); e.kind = r.get<uint32_t>(); e.value = r.get<float>();
    if (!r.ok || e.kind > 1 || e.off > pl.work_bytes || e.bytes > pl.work_bytes - e.off || (e.kind == 1 && ((e.off | e.bytes) & 3))) return S2V_EINVAL;
  }
  pl.tabs.resize(n_tab);
  for (Tab& t : pl.tabs) {
    t.off = r.get<uint64_t>(); t.bytes = r.get<uint32_t>();
    const uint32_t nrel = r.get<uint32_t>();
    if (!r.ok || t.off > pl.const_bytes || t.bytes > pl.const_bytes - t.off) return S2V_EINVAL;
    t.blob.assign(pad8(t.bytes) / 8, 0);
    if (!r.take(t.blob.data(), t.bytes) || !r.skip(pad8(t.bytes) - t.bytes)) return S2V_EINVAL;
    if (!read_relocs(r, nrel, t.bytes, pl, t.relocs)) return S2V_EINVAL;
  }
  pl.ops.resize(n_ops);
  for (Op& op : pl.ops) {
    char fn[40];
    r.take(fn, 40); fn[39] = 0;
    const uint32_t n_args = r.get<uint32_t>();
    r.get<uint32_t>();
    const Entry* e = r.ok ? find_fn(fn) : nullptr;
    if (!e || n_args != e->n_params) return S2V_EINVAL;       // unknown launcher, or a plan written against another ABI
    op.name = fn; op.fn = e->fn; op.call = e->call; op.n_params = e->n_params;
    op.args.resize(n_args);
    for (Arg& a : op.args) {
      a.kind = r.get<uint32_t>();
      const uint32_t bytes = r.get<uint32_t>();
      if (!r.ok || a.kind > kNull) return S2V_EINVAL;
      switch (a.kind) {
        case kI64: if (bytes != 8) return S2V_EINVAL; a.i = r.get<int64_t>(); break;
        case kF64: if (bytes != 8) return S2V_EINVAL; a.f = r.get<double>(); break;
        case kPtr:
          if (bytes != 16) return S2V_EINVAL;
          a.rel.at = r.get<uint32_t>(); a.rel.arena = r.get<uint32_t>(); a.rel.off = r.get<uint64_t>();
          if (!r.ok || a.rel.arena > kArenaWork || a.rel.off > (a.rel.arena == kArenaConst ? pl.const_bytes : pl.work_bytes)) return S2V_EINVAL;
          break;
        case kBlob: {
          if (bytes != sizeof(s2v_view) && bytes != sizeof(s2v_conv)) return S2V_EINVAL;     // the two structs the ABI passes by address
          a.blob_bytes = bytes;
          a.blob.assign(pad8(bytes) / 8, 0);
          if (!r.take(a.blob.data(), bytes) || !r.skip(pad8(bytes) - bytes)) return S2V_EINVAL;
          const uint32_t nrel = r.get<uint32_t>();
          r.get<uint32_t>();
          if (!read_relocs(r, nrel, bytes, pl, a.relocs)) return S2V_EINVAL;
          break;
        }
        default: if (bytes != 0) return S2V_EINVAL; break;
      }
      if (!r.ok) return S2V_EINVAL;
    }
  }
  if (!r.ok || size - r.at != pl.const_bytes) return S2V_EINVAL;
  pl.image.assign(data + r.at, data + size);
  return S2V_OK;
}

static uint8_t* base_of(const s2v_plan& pl, uint32_t arena) { return arena == kArenaConst ? pl.cdev : pl.wdev; }

__global__ void fill_f32_kernel(float* p, size_t n, float v) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = v;
}

static const Io* find_io(const s2v_plan* pl, const char* name) {
  for (const Io& e : pl->io)
    if (strcmp(e.name, name) == 0) return &e;
  return nullptr;
}
// inputs: device -> I/O slot before the replay; outputs: I/O slot -> device after it (NULL pointers are skipped)
static int copy_io(const s2v_plan* pl, const char* name, const void* src, void* dst, cudaStream_t st) {
  const Io* e = find_io(pl, name);
  if (!e) return (src || dst) ? S2V_EINVAL : S2V_OK;
  if (src) S2V_CUDA_TRY(cudaMemcpyAsync(pl->wdev + e->off, src, e->bytes, cudaMemcpyDeviceToDevice, st));
  if (dst) S2V_CUDA_TRY(cudaMemcpyAsync(dst, pl->wdev + e->off, e->bytes, cudaMemcpyDeviceToDevice, st));
  return S2V_OK;
}

}  // namespace plan
}  // namespace s2v

using namespace s2v;
using namespace s2v::plan;

extern "C" int s2v_plan_load_memory(const void* data_host, int64_t bytes, s2v_plan** out) {
  if (!data_host || bytes < 48 || !out) return S2V_EINVAL;
  s2v_plan* pl = new (std::nothrow) s2v_plan();
  if (!pl) return S2V_EINVAL;
  const int rc = parse(static_cast<const uint8_t*>(data_host), (size_t)bytes, *pl);
  if (rc != S2V_OK) { delete pl; return rc; }
  *out = pl;
  return S2V_OK;
}

extern "C" int s2v_plan_load(const char* path_host, s2v_plan** out) {
  if (!path_host || !out) return S2V_EINVAL;
  FILE* f = fopen(path_host, "rb");
  if (!f) return S2V_EINVAL;
  std::vector<uint8_t> buf;
  if (fseek(f, 0, SEEK_END) == 0) {
    const long n = ftell(f);
    if (n > 0 && fseek(f, 0, SEEK_SET) == 0) {
      buf.resize((size_t)n);
      if (fread(buf.data(), 1, (size_t)n, f) != (size_t)n) buf.clear();
    }
  }
  fclose(f);
  if (buf.empty()) return S2V_EINVAL;
  return s2v_plan_load_memory(buf.data(), (int64_t)buf.size(), out);
}

extern "C" void s2v_plan_free(s2v_plan* pl) { delete pl; }
extern "C" int64_t s2v_plan_const_bytes(const s2v_plan* pl) { return pl ? (int64_t)pl->const_bytes : -1; }
extern "C" int64_t s2v_plan_workspace_bytes(const s2v_plan* pl) { return pl ? (int64_t)pl->work_bytes : -1; }
extern "C" int s2v_plan_num_ops(const s2v_plan* pl) { return pl ? (int)pl->ops.size() : -1; }
extern "C" int s2v_plan_num_io(const s2v_plan* pl) { return pl ? (int)pl->io.size() : -1; }

extern "C" int s2v_plan_io_info(const s2v_plan* pl, int i, const char** name, int64_t* offset, int64_t* bytes, int* is_output) {
  if (!pl || i < 0 || i >= (int)pl->io.size()) return S2V_EINVAL;
  const Io& e = pl->io[(size_t)i];
  if (name) *name = e.name;
  if (offset) *offset = (int64_t)e.off;
  if (bytes) *bytes = (int64_t)e.bytes;
  if (is_output) *is_output = (int)e.is_output;
  return S2V_OK;
}

extern "C" int s2v_plan_bind(s2v_plan* pl, void* const_dev, void* workspace_dev, void* stream) {
  if (!pl || !const_dev || !workspace_dev) return S2V_EINVAL;
  if (((uintptr_t)const_dev | (uintptr_t)workspace_dev) & 255) return S2V_EINVAL;      // TMA / vector accesses assume the exporter's alignment
  cudaStream_t st = (cudaStream_t)stream;
  pl->cdev = static_cast<uint8_t*>(const_dev);
  pl->wdev = static_cast<uint8_t*>(workspace_dev);
  S2V_CUDA_TRY(cudaMemcpyAsync(pl->cdev, pl->image.data(), pl->const_bytes, cudaMemcpyHostToDevice, st));
  for (Tab& t : pl->tabs) {            // device-resident pointer tables (s2v_lin_group arrays): patch, then overwrite the stale copy
    for (const Reloc& q : t.relocs) {
      const uint64_t v = (uint64_t)(uintptr_t)(base_of(*pl, q.arena) + q.off);
      memcpy(reinterpret_cast<uint8_t*>(t.blob.data()) + q.at, &v, 8);
    }
    S2V_CUDA_TRY(cudaMemcpyAsync(pl->cdev + t.off, t.blob.data(), t.bytes, cudaMemcpyHostToDevice, st));
  }
  for (const Init& e : pl->inits) {
    if (e.bytes == 0) continue;
    if (e.kind == 0) S2V_CUDA_TRY(cudaMemsetAsync(pl->wdev + e.off, 0, e.bytes, st));
    else {
      const size_t n = e.bytes / 4;
      fill_f32_kernel<<<(unsigned)((n + 255) / 256 > 1024 ? 1024 : (n + 255) / 256), 256, 0, st>>>(reinterpret_cast<float*>(pl->wdev + e.off), n, e.value);
      S2V_CHECK_LAUNCH();
    }
  }
  for (Op& op : pl->ops)
    for (Arg& a : op.args) {
      if (a.kind == kPtr) a.ptr = base_of(*pl, a.rel.arena) + a.rel.off;
      else if (a.kind == kBlob)
        for (const Reloc& q : a.relocs) {
          const uint64_t v = (uint64_t)(uintptr_t)(base_of(*pl, q.arena) + q.off);
          memcpy(reinterpret_cast<uint8_t*>(a.blob.data()) + q.at, &v, 8);
        }
    }
  return S2V_OK;
}

extern "C" int s2v_plan_run(const s2v_plan* pl, void* stream) {
  if (!pl || !pl->cdev || !pl->wdev) return S2V_EINVAL;
  for (const Op& op : pl->ops) {
    const int rc = op.call(op.fn, op.args.data(), stream);
    if (rc != S2V_OK) return rc;
  }
  return S2V_OK;
}

extern "C" int s2v_lnet_forward(const s2v_plan* pl, const float* mel, const float* face, float* out, void* stream) {
  if (!pl || !pl->wdev || !mel || !face || !out || !find_io(pl, "mel") || !find_io(pl, "face") || !find_io(pl, "out")) return S2V_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  int rc = copy_io(pl, "mel", mel, nullptr, st);
  if (rc == S2V_OK) rc = copy_io(pl, "face", face, nullptr, st);
  if (rc == S2V_OK) rc = s2v_plan_run(pl, stream);
  if (rc == S2V_OK) rc = copy_io(pl, "out", nullptr, out, st);
  return rc;
}

extern "C" int s2v_dnet_forward(const s2v_plan* pl, const float* input_image, const float* driving_source, float* flow_field,
                                float* warp_image, float* fake_image, void* stream) {
  if (!pl || !pl->wdev || !input_image || !driving_source || !find_io(pl, "img") || !find_io(pl, "coeff")) return S2V_EINVAL;
  if (fake_image && !find_io(pl, "fake")) return S2V_EINVAL;              // a stage='warp' plan has no fake_image
  cudaStream_t st = (cudaStream_t)stream;
  int rc = copy_io(pl, "img", input_image, nullptr, st);
  if (rc == S2V_OK) rc = copy_io(pl, "coeff", driving_source, nullptr, st);
  if (rc == S2V_OK) rc = s2v_plan_run(pl, stream);
  if (rc == S2V_OK && flow_field) rc = copy_io(pl, "flow", nullptr, flow_field, st);
  if (rc == S2V_OK && warp_image) rc = copy_io(pl, "warp", nullptr, warp_image, st);
  if (rc == S2V_OK && fake_image) rc = copy_io(pl, "fake", nullptr, fake_image, st);
  return rc;
}
