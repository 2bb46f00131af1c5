// Host side of the C-ABI: handle, weight packing, workspace plan and the static launch sequence of the
// ScoreUNet forward (model/nn.py:220-242) over a batch of Markov windows.
#include <stdlib.h>

#include <algorithm>
#include <map>
#include <string>
#include <vector>

#include "c2w_b200.h"
#include "common.cuh"
#include "conv_tcgen05.cuh"
#include "kernels.cuh"

using namespace c2w;

int c2w_num_sms();

static long long g_launches = 0;  // kernels launched by this library (bench.py's gpu_launches)

namespace {

// Optional per-launch CUDA-event timing (bench.py roofline pass): events are recorded on the launch stream around
// every kernel of the forward pass and summed per class when read.
struct TimedSpan {
  int cls;  // 0 = K1 conv/GEMM (tensor cores), 1 = everything else in the forward pass
  cudaEvent_t e0, e1;
};

struct ConvW {
  bf16* w = nullptr;   // [cout_pad, taps * cin_pad], k = tap * cin_pad + c
  float* b = nullptr;  // [cout_pad]
  int cin = 0, cout = 0, cin_pad = 0, cout_pad = 0, taps = 0;
};
struct BlockW {
  int mod_off = 0;
  ConvW c1, c2;
};
struct AttnW {
  ConvW qkv, proj;
};
struct LevelW {
  int C = 0, H = 0, W = 0;
  bool attn = false;
  ConvW head, tail;
  std::vector<BlockW> desc, asc;
  std::vector<AttnW> dattn, aattn;
};

enum OpKind { OP_CONV, OP_LN, OP_ATTN };
struct Op {
  OpKind kind;
  // conv
  ConvLaunch conv;
  int pix_per_img = 0;
  bool is_final = false;
  // layernorm
  const bf16* in = nullptr;
  bf16* out = nullptr;
  int C = 0, H = 0, W = 0, up = 0, mod_off = -1;
  // attention
  int T = 0;
};

struct Plan {
  int n_max = 0;
  std::vector<Op> ops;
  bf16* xin = nullptr;
  float* h0 = nullptr;
  float* emb = nullptr;
  float* mods = nullptr;
  float* out32 = nullptr;
  size_t attn_smem = 0;
};

inline int pad64(int c) { return (c + 63) / 64 * 64; }

}  // namespace

struct c2w_handle {
  c2w_config cfg;
  int nl = 0, cin = 0, cin_pad = 0;
  std::map<std::string, std::vector<float>> raw;
  bool finalized = false;
  std::vector<void*> allocs;
  float *map0_w = nullptr, *map0_b = nullptr, *map1_w = nullptr, *map1_b = nullptr;
  float *proj_w = nullptr, *proj_b = nullptr;
  int total_mod = 0;
  std::vector<LevelW> levels;
  Plan plan;
  int sms = 0;
  bool timing = false;
  bool fuse_ln = true;  // C2W_NO_FUSE_LN=1 keeps every LayerNorm a separate kernel (A/B runs)
  std::vector<TimedSpan> spans;
  size_t spans_used = 0;
};

namespace {

struct Bump {
  uint8_t* base;
  size_t off = 0;
  explicit Bump(void* p) : base(static_cast<uint8_t*>(p)) {}
  template <typename T>
  T* take(size_t count) {
    off = (off + 1023) & ~size_t(1023);
    T* r = base ? reinterpret_cast<T*>(base + off) : nullptr;
    off += count * sizeof(T);
    return r;
  }
};

int dev_upload(c2w_handle* h, const void* src, size_t bytes, void** out) {
  void* d = nullptr;
  C2W_CUDA(cudaMalloc(&d, bytes));
  h->allocs.push_back(d);
  C2W_CUDA(cudaMemcpy(d, src, bytes, cudaMemcpyHostToDevice));
  *out = d;
  return C2W_OK;
}

int need(c2w_handle* h, const std::string& name, size_t numel, const std::vector<float>** out) {
  auto it = h->raw.find(name);
  if (it == h->raw.end()) return fail(C2W_ERR_MISSING, "weight '%s' was never loaded", name.c_str());
  if (it->second.size() != numel)
    return fail(C2W_ERR_INVALID, "weight '%s': expected %zu elements, got %zu", name.c_str(), numel, it->second.size());
  *out = &it->second;
  return C2W_OK;
}

// OIHW (or OI1 for Conv1d) fp32 -> bf16 [cout_pad, taps*cin_pad] with k = (r*kw + s)*cin_pad + c; bias zero-padded.
int pack_conv(c2w_handle* h, const std::string& prefix, int cout, int cin, int taps, ConvW* cw) {
  const std::vector<float>*w, *b;
  int rc = need(h, prefix + ".weight", static_cast<size_t>(cout) * cin * taps, &w);
  if (rc) return rc;
  rc = need(h, prefix + ".bias", cout, &b);
  if (rc) return rc;
  cw->cin = cin;
  cw->cout = cout;
  cw->taps = taps;
  cw->cin_pad = pad64(cin);
  cw->cout_pad = pad64(cout);
  const size_t K = static_cast<size_t>(taps) * cw->cin_pad;
  std::vector<bf16> packed(static_cast<size_t>(cw->cout_pad) * K, __float2bfloat16_rn(0.f));
  for (int o = 0; o < cout; ++o)
    for (int c = 0; c < cin; ++c)
      for (int t = 0; t < taps; ++t)
        packed[o * K + static_cast<size_t>(t) * cw->cin_pad + c] =
            __float2bfloat16_rn((*w)[(static_cast<size_t>(o) * cin + c) * taps + t]);
  std::vector<float> bias(cw->cout_pad, 0.f);
  for (int o = 0; o < cout; ++o) bias[o] = (*b)[o];
  rc = dev_upload(h, packed.data(), packed.size() * sizeof(bf16), reinterpret_cast<void**>(&cw->w));
  if (rc) return rc;
  return dev_upload(h, bias.data(), bias.size() * sizeof(float), reinterpret_cast<void**>(&cw->b));
}

int upload_named(c2w_handle* h, const std::string& name, size_t numel, float** out) {
  const std::vector<float>* v;
  int rc = need(h, name, numel, &v);
  if (rc) return rc;
  return dev_upload(h, v->data(), numel * sizeof(float), reinterpret_cast<void**>(out));
}

template <int C>
void launch_ln_c(const bf16* x, const float* mod, bf16* out, long long npix, int H, int W, int up, int sms,
                 cudaStream_t st) {
  const int threads = 256;
  long long blocks = (npix + 7) / 8;
  const long long cap = static_cast<long long>(sms) * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  channel_layernorm_kernel<C><<<static_cast<int>(blocks), threads, 0, st>>>(x, mod, out, npix, H, W, up, 1e-5f);
}

int launch_ln(const bf16* x, const float* mod, bf16* out, long long npix, int C, int H, int W, int up, int sms,
              cudaStream_t st) {
  switch (C) {
    case 64: launch_ln_c<64>(x, mod, out, npix, H, W, up, sms, st); break;
    case 128: launch_ln_c<128>(x, mod, out, npix, H, W, up, sms, st); break;
    case 192: launch_ln_c<192>(x, mod, out, npix, H, W, up, sms, st); break;
    case 256: launch_ln_c<256>(x, mod, out, npix, H, W, up, sms, st); break;
    case 320: launch_ln_c<320>(x, mod, out, npix, H, W, up, sms, st); break;
    case 384: launch_ln_c<384>(x, mod, out, npix, H, W, up, sms, st); break;
    case 448: launch_ln_c<448>(x, mod, out, npix, H, W, up, sms, st); break;
    case 512: launch_ln_c<512>(x, mod, out, npix, H, W, up, sms, st); break;
    default: return fail(C2W_ERR_INVALID, "channel LayerNorm: unsupported C=%d (multiple of 64, <= 512)", C);
  }
  C2W_CUDA(cudaGetLastError());
  return C2W_OK;
}

int launch_attention(const bf16* qkv, bf16* out, int n, int T, int C, cudaStream_t st) {
  C2W_REQUIRE(T % 4 == 0 && C % 8 == 0, "attention: T %% 4 and C %% 8 must be 0 (T=%d C=%d)", T, C);
  // 64-query CTAs when they already fill the machine twice over, else 16-query CTAs (4x the parallelism)
  const int sms = c2w_num_sms();
  const bool big = static_cast<long long>(n) * ((T + 63) / 64) >= 2LL * sms;
  const int qb = big ? 64 : 16;
  const size_t smem = attention_smem_bytes(T, C, qb);
  C2W_REQUIRE(smem <= static_cast<size_t>(kSmemLimit), "attention: T=%d C=%d needs %zu B of shared memory", T, C, smem);
  static size_t configured[2] = {0, 0};
  const float scale2 = 1.0f / sqrtf(static_cast<float>(C));
  dim3 grid((T + qb - 1) / qb, n);
  if (big) {
    if (smem > configured[0]) {
      C2W_CUDA(cudaFuncSetAttribute(attention_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
      configured[0] = smem;
    }
    attention_kernel<64><<<grid, kAttnThreads, smem, st>>>(qkv, out, T, C, scale2);
  } else {
    if (smem > configured[1]) {
      C2W_CUDA(cudaFuncSetAttribute(attention_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
      configured[1] = smem;
    }
    attention_kernel<16><<<grid, kAttnThreads, smem, st>>>(qkv, out, T, C, scale2);
  }
  C2W_CUDA(cudaGetLastError());
  return C2W_OK;
}

int grid_for(long long items, int threads, int sms) {
  long long b = (items + threads - 1) / threads;
  const long long cap = static_cast<long long>(sms) * 16;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return static_cast<int>(b);
}

int run_modulation(c2w_handle* h, float t, float* h0, float* emb, float* mods, cudaStream_t st) {
  const int E = h->cfg.embedding_dim, nf = h->cfg.noise_features;
  time_embed_kernel<<<1, 256, nf * sizeof(float), st>>>(t, h->map0_w, h->map0_b, h0, E, nf);
  matvec_kernel<<<ceil_div(static_cast<long long>(E) * 32, 256), 256, 0, st>>>(h->map1_w, h->map1_b, h0, emb, E, E, 1);
  matvec_kernel<<<ceil_div(static_cast<long long>(h->total_mod) * 32, 256), 256, 0, st>>>(h->proj_w, h->proj_b, emb, mods,
                                                                                         h->total_mod, E, 0);
  g_launches += 3;
  C2W_CUDA(cudaGetLastError());
  return C2W_OK;
}

// Lays the workspace out and (if base != null) builds every launch of the forward pass over it.
int build_plan(c2w_handle* h, int n, void* base, size_t* bytes_out) {
  Plan& P = h->plan;
  const bool real = base != nullptr;
  Bump B(base);
  const int nl = h->nl;
  const long long HW0 = static_cast<long long>(h->cfg.height) * h->cfg.width;
  if (real) {
    P.ops.clear();
    P.n_max = n;
  }
  bf16* xin = B.take<bf16>(n * HW0 * h->cin_pad);
  float* h0 = B.take<float>(h->cfg.embedding_dim);
  float* emb = B.take<float>(h->cfg.embedding_dim);
  float* mods = B.take<float>(h->total_mod);
  float* out32 = B.take<float>(n * HW0 * h->levels[0].tail.cout_pad);
  std::vector<bf16*> xs(nl), as(nl), hs(nl);
  size_t up_elems = 0, qkv_elems = 0, att_elems = 0;
  for (int l = 0; l < nl; ++l) {
    const LevelW& L = h->levels[l];
    const size_t e = static_cast<size_t>(n) * L.H * L.W * L.C;
    xs[l] = B.take<bf16>(e);
    as[l] = B.take<bf16>(e);
    hs[l] = B.take<bf16>(e);
    if (l > 0) {
      const LevelW& U = h->levels[l - 1];
      up_elems = std::max(up_elems, static_cast<size_t>(n) * U.H * U.W * L.C);
    }
    if (L.attn) {
      qkv_elems = std::max(qkv_elems, static_cast<size_t>(n) * L.H * L.W * 3 * L.C);
      att_elems = std::max(att_elems, e);
    }
  }
  bf16* up = B.take<bf16>(up_elems);
  bf16* qkv = B.take<bf16>(qkv_elems);
  bf16* att = B.take<bf16>(att_elems);
  *bytes_out = B.off + 1024;
  if (!real) return C2W_OK;
  P.xin = xin;
  P.h0 = h0;
  P.emb = emb;
  P.mods = mods;
  P.out32 = out32;

  // H, W: INPUT image size; stride 2 halves it (heads of levels > 0, model/nn.py:169-176)
  auto add_conv = [&](bool c3, const bf16* in, int H, int W, int cin, const ConvW& w, int mode, const bf16* res,
                      bf16* out, bool is_final, int stride = 1) -> int {
    Op op;
    op.kind = OP_CONV;
    const int bn = conv_pick_bn(w.cout_pad);
    if (!conv_launch_init(&op.conv, c3, in, n, H, W, cin, w.w, w.cout_pad, bn, h->sms, stride))
      return fail(C2W_ERR_INVALID, "cannot build conv launch (H=%d W=%d cin=%d cout=%d stride=%d)", H, W, cin,
                  w.cout_pad, stride);
    op.conv.p.mode = mode;
    op.conv.p.bias = w.b;
    if (mode == EPI_BIAS_RES && res != out)
      return fail(C2W_ERR_INVALID, "residual convs accumulate in place (res must be out)");
    if (out && !conv_launch_set_out(&op.conv, out)) return fail(C2W_ERR_CUDA, "cannot encode the output tensor map");
    op.pix_per_img = (H / stride) * (W / stride);
    op.is_final = is_final;
    P.ops.push_back(op);
    return C2W_OK;
  };
  // Channel LayerNorm of `in` (+ modulation) -> `out`.  When `in` was just produced by a conv whose single N tile
  // holds all C channels of a pixel, the normalisation is done in that conv's epilogue (no extra pass over HBM).
  auto add_ln = [&](const bf16* in, bf16* out, int C, int H, int W, int upf, int mod_off) {
    if (h->fuse_ln && !P.ops.empty()) {
      Op& prev = P.ops.back();
      if (prev.kind == OP_CONV && !prev.is_final && prev.conv.out_ptr == in && prev.conv.cout_pad == C &&
          conv_launch_can_ln(&prev.conv, upf) &&
          conv_launch_set_ln(&prev.conv, out, mod_off >= 0 ? mods + mod_off : nullptr, upf))
        return;
    }
    Op op;
    op.kind = OP_LN;
    op.in = in;
    op.out = out;
    op.C = C;
    op.H = H;
    op.W = W;
    op.up = upf;
    op.mod_off = mod_off;
    P.ops.push_back(op);
  };
  auto add_blocks = [&](int l, const std::vector<BlockW>& blocks, const std::vector<AttnW>& attns) -> int {
    const LevelW& L = h->levels[l];
    for (size_t b = 0; b < blocks.size(); ++b) {
      const BlockW& bw = blocks[b];
      // x + conv2(SiLU(conv1(LN(x + proj(emb)))))    model/nn.py:27-28,151-158
      add_ln(xs[l], as[l], L.C, L.H, L.W, 0, bw.mod_off);
      int rc = add_conv(true, as[l], L.H, L.W, L.C, bw.c1, EPI_BIAS_SILU, nullptr, hs[l], false);
      if (rc) return rc;
      rc = add_conv(true, hs[l], L.H, L.W, L.C, bw.c2, EPI_BIAS_RES, xs[l], xs[l], false);
      if (rc) return rc;
      if (L.attn) {
        // x + proj(attn(qkv(LN(x))))    model/nn.py:50-60
        const AttnW& aw = attns[b];
        add_ln(xs[l], as[l], L.C, L.H, L.W, 0, -1);
        rc = add_conv(false, as[l], L.H, L.W, L.C, aw.qkv, EPI_BIAS, nullptr, qkv, false);
        if (rc) return rc;
        Op op;
        op.kind = OP_ATTN;
        op.in = qkv;
        op.out = att;
        op.T = L.H * L.W;
        op.C = L.C;
        P.ops.push_back(op);
        P.attn_smem = std::max(P.attn_smem, attention_smem_bytes(op.T, op.C));
        if (P.attn_smem > static_cast<size_t>(kSmemLimit))
          return fail(C2W_ERR_INVALID, "attention at level %d (T=%d, C=%d) exceeds shared memory", l, op.T, op.C);
        rc = add_conv(false, att, L.H, L.W, L.C, aw.proj, EPI_BIAS_RES, xs[l], xs[l], false);
        if (rc) return rc;
      }
    }
    return C2W_OK;
  };

  // descent (model/nn.py:223-229)
  for (int l = 0; l < nl; ++l) {
    const LevelW& L = h->levels[l];
    int rc;
    if (l == 0) {
      rc = add_conv(true, xin, L.H, L.W, h->cin_pad, L.head, EPI_BIAS, nullptr, xs[0], false);
    } else {
      const LevelW& U = h->levels[l - 1];
      rc = add_conv(true, xs[l - 1], U.H, U.W, U.C, L.head, EPI_BIAS, nullptr, xs[l], false, 2);
    }
    if (rc) return rc;
    rc = add_blocks(l, L.desc, L.dattn);
    if (rc) return rc;
  }
  // ascent (model/nn.py:233-240): blocks, then tail = LN -> nearest x2 -> conv, + skip (in place into the skip)
  for (int l = nl - 1; l >= 0; --l) {
    const LevelW& L = h->levels[l];
    int rc = add_blocks(l, L.asc, L.aattn);
    if (rc) return rc;
    if (l > 0) {
      const LevelW& U = h->levels[l - 1];
      add_ln(xs[l], up, L.C, L.H, L.W, 1, -1);
      rc = add_conv(true, up, U.H, U.W, L.C, L.tail, EPI_BIAS_RES, xs[l - 1], xs[l - 1], false);
    } else {
      rc = add_conv(true, xs[0], L.H, L.W, L.C, L.tail, EPI_F32, nullptr, nullptr, true);
    }
    if (rc) return rc;
  }
  return C2W_OK;
}

struct SpanGuard {
  c2w_handle* h;
  cudaStream_t st;
  TimedSpan* sp = nullptr;
  SpanGuard(c2w_handle* h_, int cls, cudaStream_t st_) : h(h_), st(st_) {
    ++g_launches;
    if (!h->timing) return;
    if (h->spans_used == h->spans.size()) {
      TimedSpan t;
      t.cls = cls;
      cudaEventCreate(&t.e0);
      cudaEventCreate(&t.e1);
      h->spans.push_back(t);
    }
    sp = &h->spans[h->spans_used++];
    sp->cls = cls;
    cudaEventRecord(sp->e0, st);
  }
  ~SpanGuard() {
    if (sp) cudaEventRecord(sp->e1, st);
  }
};

struct FinalSpec {
  int mode;  // EPI_F32 or EPI_COMPOSE
  float* eps = nullptr;
  int order_k = 0, win_first = 0, win_last_global = 0, frame_base = 0;
};

// Runs the forward pass on the first nn windows of plan.xin.
int run_plan(c2w_handle* h, int nn, const FinalSpec& fs, cudaStream_t st) {
  Plan& P = h->plan;
  for (Op& op : P.ops) {
    switch (op.kind) {
      case OP_CONV: {
        ConvLaunch L = op.conv;
        const long long m_total = static_cast<long long>(nn) * op.pix_per_img;
        L.p.m_total = static_cast<int>(m_total);
        L.p.num_m_tiles = ceil_div(m_total, kBlockM);
        conv_set_grid(&L, h->sms);
        if (op.is_final) {
          L.p.mode = fs.mode;
          L.p.out_f32 = P.out32;
          L.p.eps = fs.eps;
          L.p.hw = op.pix_per_img;
          L.p.order_k = fs.order_k;
          L.p.win_first = fs.win_first;
          L.p.win_last_global = fs.win_last_global;
          L.p.frame_base = fs.frame_base;
        }
        {
          SpanGuard sg(h, 0, st);
          C2W_CUDA(conv_launch(L, st));
        }
        break;
      }
      case OP_LN: {
        SpanGuard sg(h, 1, st);
        int rc = launch_ln(op.in, op.mod_off >= 0 ? P.mods + op.mod_off : nullptr, op.out,
                           static_cast<long long>(nn) * op.H * op.W, op.C, op.H, op.W, op.up, h->sms, st);
        if (rc) return rc;
        break;
      }
      case OP_ATTN: {
        SpanGuard sg(h, 1, st);
        int rc = launch_attention(op.in, op.out, nn, op.T, op.C, st);
        if (rc) return rc;
        break;
      }
    }
  }
  return C2W_OK;
}

}  // namespace

extern "C" {

int c2w_abi_version(void) { return 1; }
const char* c2w_last_error(void) { return error_slot(); }

int c2w_create(const c2w_config* cfg, c2w_handle** out) {
  C2W_REQUIRE(cfg && out, "c2w_create: null argument");
  C2W_REQUIRE(cfg->n_levels >= 1 && cfg->n_levels <= C2W_MAX_LEVELS, "n_levels=%d out of range", cfg->n_levels);
  C2W_REQUIRE(cfg->frame_channels >= 1 && cfg->window >= 1 && (cfg->window % 2) == 1, "window must be odd, got %d",
              cfg->window);
  C2W_REQUIRE(cfg->embedding_dim % 4 == 0 && cfg->noise_features % 2 == 0 && cfg->noise_features <= 256,
              "embedding_dim %% 4 == 0 and even noise_features <= 256 required");
  int H = cfg->height, W = cfg->width;
  for (int l = 0; l < cfg->n_levels; ++l) {
    C2W_REQUIRE(cfg->hidden_channels[l] % 64 == 0 && cfg->hidden_channels[l] <= 512,
                "hidden_channels[%d]=%d: must be a multiple of 64 and <= 512", l, cfg->hidden_channels[l]);
    C2W_REQUIRE(W >= 8 && W <= 128 && 128 % W == 0 && H >= 1, "level %d: width %d must be a power of two in [8,128]", l, W);
    const int th = std::min(H, 128 / W);
    C2W_REQUIRE(H % th == 0 && 128 % (W * th) == 0, "level %d: %dx%d does not tile into 128-pixel row blocks", l, H, W);
    if (l + 1 < cfg->n_levels) {
      C2W_REQUIRE(H % 2 == 0 && W % 2 == 0, "level %d: %dx%d not divisible by the stride", l, H, W);
      H /= 2;
      W /= 2;
    }
  }
  int sms = c2w_num_sms();
  C2W_REQUIRE(sms > 0, "c2w_create: no CUDA device available (this library has no CPU path)");
  c2w_handle* h = new c2w_handle();
  h->cfg = *cfg;
  h->nl = cfg->n_levels;
  h->cin = cfg->frame_channels * cfg->window;
  h->cin_pad = pad64(h->cin);
  h->sms = sms;
  {
    const char* e = getenv("C2W_NO_FUSE_LN");
    h->fuse_ln = !(e && e[0] == '1');
  }
  *out = h;
  return C2W_OK;
}

void c2w_destroy(c2w_handle* h) {
  if (!h) return;
  for (void* p : h->allocs) cudaFree(p);
  delete h;
}

int c2w_load_weight(c2w_handle* h, const char* name, const float* host_data, int64_t numel) {
  C2W_REQUIRE(h && name && host_data && numel > 0, "c2w_load_weight: bad argument");
  if (h->finalized) return fail(C2W_ERR_STATE, "weights already finalised");
  h->raw[name].assign(host_data, host_data + numel);
  return C2W_OK;
}

int c2w_finalize_weights(c2w_handle* h) {
  C2W_REQUIRE(h, "null handle");
  if (h->finalized) return fail(C2W_ERR_STATE, "weights already finalised");
  const c2w_config& c = h->cfg;
  const int E = c.embedding_dim, nl = h->nl;
  int rc;
  if ((rc = upload_named(h, "map_layer0.weight", static_cast<size_t>(E) * c.noise_features, &h->map0_w))) return rc;
  if ((rc = upload_named(h, "map_layer0.bias", E, &h->map0_b))) return rc;
  if ((rc = upload_named(h, "map_layer1.weight", static_cast<size_t>(E) * E, &h->map1_w))) return rc;
  if ((rc = upload_named(h, "map_layer1.bias", E, &h->map1_b))) return rc;
  h->levels.assign(nl, LevelW());
  std::vector<float> proj_w, proj_b;
  int H = c.height, W = c.width;
  for (int l = 0; l < nl; ++l) {
    LevelW& L = h->levels[l];
    const int rev = nl - 1 - l;
    L.C = c.hidden_channels[l];
    L.H = H;
    L.W = W;
    L.attn = (c.attention_mask >> l) & 1;
    char buf[128];
    if (l == 0) {
      if ((rc = pack_conv(h, "unet.heads.0", L.C, h->cin, 9, &L.head))) return rc;
      snprintf(buf, sizeof buf, "unet.tails.%d", rev);
      if ((rc = pack_conv(h, buf, h->cin, L.C, 9, &L.tail))) return rc;
    } else {
      const int Cu = c.hidden_channels[l - 1];
      snprintf(buf, sizeof buf, "unet.heads.%d.0", l);
      if ((rc = pack_conv(h, buf, L.C, Cu, 9, &L.head))) return rc;
      snprintf(buf, sizeof buf, "unet.tails.%d.2", rev);
      if ((rc = pack_conv(h, buf, Cu, L.C, 9, &L.tail))) return rc;
    }
    const int step = L.attn ? 2 : 1;
    for (int side = 0; side < 2; ++side) {
      std::vector<BlockW>& blocks = side == 0 ? L.desc : L.asc;
      std::vector<AttnW>& attns = side == 0 ? L.dattn : L.aattn;
      for (int b = 0; b < c.hidden_blocks[l]; ++b) {
        snprintf(buf, sizeof buf, "unet.%s.%d.%d", side == 0 ? "descent" : "ascent", side == 0 ? l : rev, b * step);
        const std::string p(buf);
        BlockW bw;
        const std::vector<float>*pw, *pb;
        if ((rc = need(h, p + ".project.0.weight", static_cast<size_t>(L.C) * E, &pw))) return rc;
        if ((rc = need(h, p + ".project.0.bias", L.C, &pb))) return rc;
        bw.mod_off = static_cast<int>(proj_b.size());
        proj_w.insert(proj_w.end(), pw->begin(), pw->end());
        proj_b.insert(proj_b.end(), pb->begin(), pb->end());
        if ((rc = pack_conv(h, p + ".residue.1", L.C, L.C, 9, &bw.c1))) return rc;
        if ((rc = pack_conv(h, p + ".residue.3", L.C, L.C, 9, &bw.c2))) return rc;
        blocks.push_back(bw);
        if (L.attn) {
          snprintf(buf, sizeof buf, "unet.%s.%d.%d", side == 0 ? "descent" : "ascent", side == 0 ? l : rev, b * step + 1);
          const std::string q(buf);
          AttnW aw;
          if ((rc = pack_conv(h, q + ".qkv", 3 * L.C, L.C, 1, &aw.qkv))) return rc;
          if ((rc = pack_conv(h, q + ".proj_out", L.C, L.C, 1, &aw.proj))) return rc;
          attns.push_back(aw);
        }
      }
    }
    H /= 2;
    W /= 2;
  }
  h->total_mod = static_cast<int>(proj_b.size());
  if ((rc = dev_upload(h, proj_w.data(), proj_w.size() * sizeof(float), reinterpret_cast<void**>(&h->proj_w)))) return rc;
  if ((rc = dev_upload(h, proj_b.data(), proj_b.size() * sizeof(float), reinterpret_cast<void**>(&h->proj_b)))) return rc;
  h->raw.clear();
  h->finalized = true;
  return C2W_OK;
}

int c2w_total_mod_channels(c2w_handle* h) { return h ? h->total_mod : 0; }

int64_t c2w_launch_count(void) { return g_launches; }

int c2w_set_timing(c2w_handle* h, int enable) {
  C2W_REQUIRE(h, "null handle");
  h->timing = enable != 0;
  h->spans_used = 0;
  return C2W_OK;
}

// Sums the event-timed spans recorded since the last read: ms[0] / n[0] = K1 conv/GEMM launches, ms[1] / n[1] = the
// other forward-pass kernels.  Synchronises on the recorded events.
int c2w_timing_read(c2w_handle* h, double* ms, int64_t* n) {
  C2W_REQUIRE(h && ms && n, "c2w_timing_read: bad argument");
  ms[0] = ms[1] = 0.0;
  n[0] = n[1] = 0;
  for (size_t i = 0; i < h->spans_used; ++i) {
    TimedSpan& s = h->spans[i];
    C2W_CUDA(cudaEventSynchronize(s.e1));
    float t = 0.f;
    C2W_CUDA(cudaEventElapsedTime(&t, s.e0, s.e1));
    ms[s.cls] += t;
    n[s.cls] += 1;
  }
  h->spans_used = 0;
  return C2W_OK;
}

int64_t c2w_workspace_bytes(c2w_handle* h, int32_t max_windows) {
  if (!h || !h->finalized || max_windows < 1) {
    fail(C2W_ERR_STATE, "c2w_workspace_bytes: finalise the weights first");
    return -1;
  }
  size_t bytes = 0;
  if (build_plan(h, max_windows, nullptr, &bytes)) return -1;
  return static_cast<int64_t>(bytes);
}

int c2w_bind_workspace(c2w_handle* h, int32_t max_windows, void* dev_ptr, int64_t bytes) {
  C2W_REQUIRE(h && dev_ptr && max_windows >= 1, "c2w_bind_workspace: bad argument");
  if (!h->finalized) return fail(C2W_ERR_STATE, "finalise the weights first");
  size_t need_bytes = 0;
  int rc = build_plan(h, max_windows, nullptr, &need_bytes);
  if (rc) return rc;
  C2W_REQUIRE(static_cast<size_t>(bytes) >= need_bytes, "workspace too small: %lld < %zu", (long long)bytes, need_bytes);
  void* aligned = reinterpret_cast<void*>((reinterpret_cast<uintptr_t>(dev_ptr) + 1023) & ~uintptr_t(1023));
  return build_plan(h, max_windows, aligned, &need_bytes);
}

int c2w_op_modulation(c2w_handle* h, float t, float* emb_out, float* mods_out, void* stream) {
  C2W_REQUIRE(h && h->finalized && h->plan.n_max > 0, "c2w_op_modulation: bind a workspace first");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int rc = run_modulation(h, t, h->plan.h0, h->plan.emb, h->plan.mods, st);
  if (rc) return rc;
  if (emb_out)
    C2W_CUDA(cudaMemcpyAsync(emb_out, h->plan.emb, h->cfg.embedding_dim * sizeof(float), cudaMemcpyDeviceToDevice, st));
  if (mods_out)
    C2W_CUDA(cudaMemcpyAsync(mods_out, h->plan.mods, h->total_mod * sizeof(float), cudaMemcpyDeviceToDevice, st));
  return C2W_OK;
}

int c2w_unet_forward(c2w_handle* h, const float* x_nchw, int32_t n, float t, float* out_nchw, void* stream) {
  C2W_REQUIRE(h && x_nchw && out_nchw && n >= 1, "c2w_unet_forward: bad argument");
  if (h->plan.n_max < 1) return fail(C2W_ERR_STATE, "bind a workspace first");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  Plan& P = h->plan;
  const int hw = h->cfg.height * h->cfg.width;
  const int cout_pad = h->levels[0].tail.cout_pad;
  int rc = run_modulation(h, t, P.h0, P.emb, P.mods, st);
  if (rc) return rc;
  FinalSpec fs;
  fs.mode = EPI_F32;
  for (int i0 = 0; i0 < n; i0 += P.n_max) {
    const int nn = std::min(P.n_max, n - i0);
    dim3 blk(32, 8);
    dim3 g1(ceil_div(hw, 32), ceil_div(h->cin_pad, 32), nn);
    nchw_to_nhwc_bf16_kernel<<<g1, blk, 0, st>>>(x_nchw + static_cast<size_t>(i0) * h->cin * hw, P.xin, h->cin, hw,
                                                 h->cin_pad);
    C2W_CUDA(cudaGetLastError());
    rc = run_plan(h, nn, fs, st);
    if (rc) return rc;
    dim3 g2(ceil_div(hw, 32), ceil_div(cout_pad, 32), nn);
    nhwc_f32_to_nchw_kernel<<<g2, blk, 0, st>>>(P.out32, out_nchw + static_cast<size_t>(i0) * h->cin * hw, h->cin, hw,
                                                cout_pad);
    C2W_CUDA(cudaGetLastError());
  }
  return C2W_OK;
}

int c2w_window_score(c2w_handle* h, const float* traj, int32_t n_frames_local, int32_t frame_global0,
                     int32_t win_first, int32_t n_win, int32_t n_win_global, float t, float* eps, void* stream) {
  C2W_REQUIRE(h && traj && eps, "c2w_window_score: bad argument");
  if (h->plan.n_max < 1) return fail(C2W_ERR_STATE, "bind a workspace first");
  const int w = h->cfg.window, k = w / 2, C = h->cfg.frame_channels;
  C2W_REQUIRE(n_win >= 1 && win_first >= 0 && win_first + n_win <= n_win_global, "window range [%d,%d) outside [0,%d)",
              win_first, win_first + n_win, n_win_global);
  C2W_REQUIRE(win_first >= frame_global0 && win_first + n_win - 1 + w <= frame_global0 + n_frames_local,
              "windows [%d,%d) need frames outside the local range [%d,%d)", win_first, win_first + n_win,
              frame_global0, frame_global0 + n_frames_local);
  C2W_REQUIRE(C == 4, "fused compose supports 4 variables per frame (got %d)", C);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  Plan& P = h->plan;
  const int hw = h->cfg.height * h->cfg.width;
  int rc = run_modulation(h, t, P.h0, P.emb, P.mods, st);
  if (rc) return rc;
  for (int c0 = 0; c0 < n_win; c0 += P.n_max) {
    const int nn = std::min(P.n_max, n_win - c0);
    const int j0 = win_first + c0;
    const long long items = static_cast<long long>(nn) * hw * (h->cin_pad / 8);
    gather_windows_kernel<<<grid_for(items, 256, h->sms), 256, 0, st>>>(traj, P.xin, nn, hw, C, w * C, h->cin_pad,
                                                                        j0 - frame_global0);
    ++g_launches;
    C2W_CUDA(cudaGetLastError());
    FinalSpec fs;
    fs.mode = EPI_COMPOSE;
    fs.eps = eps;
    fs.order_k = k;
    fs.win_first = j0;
    fs.win_last_global = n_win_global - 1;
    fs.frame_base = frame_global0;
    rc = run_plan(h, nn, fs, st);
    if (rc) return rc;
  }
  return C2W_OK;
}

int c2w_traj_pack(const float* nchw, float* fhwc, int64_t frames, int32_t C, int32_t hw, void* stream) {
  C2W_REQUIRE(nchw && fhwc && frames >= 1, "c2w_traj_pack: bad argument");
  nchw_to_fhwc_kernel<<<grid_for(frames * hw, 256, c2w_num_sms()), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      nchw, fhwc, frames, C, hw);
  C2W_CUDA(cudaGetLastError());
  return C2W_OK;
}
int c2w_traj_unpack(const float* fhwc, float* nchw, int64_t frames, int32_t C, int32_t hw, void* stream) {
  C2W_REQUIRE(nchw && fhwc && frames >= 1, "c2w_traj_unpack: bad argument");
  fhwc_to_nchw_kernel<<<grid_for(frames * hw, 256, c2w_num_sms()), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      fhwc, nchw, frames, C, hw);
  C2W_CUDA(cudaGetLastError());
  return C2W_OK;
}

int c2w_guided_step(const c2w_guide* g, void* stream) {
  C2W_REQUIRE(g && g->x && g->eps && g->nan_flag, "c2w_guided_step: bad argument");
  C2W_REQUIRE(g->s_step >= 1 && g->H % g->s_step == 0 && g->W % g->s_step == 0 && g->W / g->s_step <= 32,
              "s_step=%d must divide %dx%d with at most 32 tiles per row", g->s_step, g->H, g->W);
  C2W_REQUIRE(g->mode == 0 || (g->eps_out && g->partials), "mode 1 needs eps_out and partials");
  C2W_REQUIRE(g->own_n >= 1 && g->t_step >= 1, "own_n and t_step must be positive");
  GuideParams p;
  p.x = g->x;
  p.eps = g->eps;
  p.eps_out = g->eps_out;
  p.y = g->y;
  for (int i = 0; i < 4; ++i) {
    p.std2[i] = g->std2[i];
    p.gamma[i] = g->gamma[i];
  }
  p.mu = g->mu;
  p.sigma = g->sigma;
  p.mu_next = g->mu_next;
  p.sigma_next = g->sigma_next;
  p.t_step = g->t_step;
  p.s_step = g->s_step;
  p.H = g->H;
  p.W = g->W;
  p.frame_global0 = g->frame_global0;
  p.own_lo = g->own_lo;
  p.mode = g->mode;
  p.partials = g->partials;
  p.nan_flag = g->nan_flag;
  dim3 grid(g->H / g->s_step, g->own_n);
  guided_step_kernel<<<grid, 32 * (g->W / g->s_step), 0, static_cast<cudaStream_t>(stream)>>>(p);
  ++g_launches;
  C2W_CUDA(cudaGetLastError());
  return C2W_OK;
}

int c2w_reduce_partials(const float* partials, int32_t n, double* sumsq, void* stream) {
  C2W_REQUIRE(partials && sumsq && n >= 1, "c2w_reduce_partials: bad argument");
  reduce_partials_kernel<<<1, 256, 0, static_cast<cudaStream_t>(stream)>>>(partials, n, sumsq);
  ++g_launches;
  C2W_CUDA(cudaGetLastError());
  return C2W_OK;
}

int c2w_corrector_update(float* x, const float* eps, const float* z, const double* sumsq, double count, float tau,
                         float sigma_next, int64_t pix0_global, int64_t npix, uint64_t seed, uint32_t step_id,
                         int32_t* nan_flag, void* stream) {
  C2W_REQUIRE(x && eps && sumsq && nan_flag && npix >= 1 && count > 0, "c2w_corrector_update: bad argument");
  corrector_update_kernel<<<grid_for(npix, 256, c2w_num_sms()), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      x, eps, z, sumsq, count, tau, sigma_next, pix0_global, npix, seed, step_id, nan_flag);
  ++g_launches;
  C2W_CUDA(cudaGetLastError());
  return C2W_OK;
}

// ---------------------------------------------------------------------------------------------- op-level hooks
int c2w_op_layernorm(const void* x, const float* mod, void* out, int64_t npix, int C, int H, int W, int upsample,
                     void* stream) {
  C2W_REQUIRE(x && out && npix >= 1, "c2w_op_layernorm: bad argument");
  return launch_ln(static_cast<const bf16*>(x), mod, static_cast<bf16*>(out), npix, C, H, W, upsample, c2w_num_sms(),
                   static_cast<cudaStream_t>(stream));
}
int c2w_op_attention(const void* qkv, void* out, int n, int T, int C, void* stream) {
  C2W_REQUIRE(qkv && out && n >= 1, "c2w_op_attention: bad argument");
  return launch_attention(static_cast<const bf16*>(qkv), static_cast<bf16*>(out), n, T, C,
                          static_cast<cudaStream_t>(stream));
}
int c2w_op_gather_windows(const float* traj, void* out, int n, int hw, int C, int window, int cin_pad, int frame0,
                          void* stream) {
  C2W_REQUIRE(traj && out && n >= 1 && cin_pad % 8 == 0 && cin_pad >= C * window, "c2w_op_gather_windows: bad argument");
  const long long items = static_cast<long long>(n) * hw * (cin_pad / 8);
  gather_windows_kernel<<<grid_for(items, 256, c2w_num_sms()), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      traj, static_cast<bf16*>(out), n, hw, C, window * C, cin_pad, frame0);
  C2W_CUDA(cudaGetLastError());
  return C2W_OK;
}

}  // extern "C"
