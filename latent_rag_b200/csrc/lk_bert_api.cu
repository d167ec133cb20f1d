// C ABI of the sentence encoder (include/latentknn.h, lk_bert_*): the forward of the SBERT model
// the reference embeds its corpus and queries with (SentenceTransformer("all-MiniLM-L6-v2").encode,
// retrieval/embedder.py:35-40; main.py builds the corpus embeddings through it): a BERT encoder,
// masked mean pooling and L2 normalisation.  Tokenisation stays on the host, outside this library.
#include <cstring>
#include <new>
#include <vector>

#include "lk_common.cuh"
#include "lk_host.cuh"

namespace lk {
void ae_umma_weight_slabs(const float* w, int rows_out, int k_in, int slab_rows, std::vector<unsigned char>* out,
                          bool plane_major_over_kb_only);
}

using namespace lk;

namespace {

struct BertLayer {
  unsigned char *wqkv = nullptr, *wo = nullptr, *w1 = nullptr, *w2 = nullptr;  // split-bf16 slab planes
  float *bqkv = nullptr, *bo = nullptr, *b1 = nullptr, *b2 = nullptr;
  float *ln1_g = nullptr, *ln1_b = nullptr, *ln2_g = nullptr, *ln2_b = nullptr;
};

constexpr int64_t kChunkTokens = 1 << 16;  // tokens per pass of the layer stack

}  // namespace

struct lk_bert {
  int device = 0, sm_count = 0;
  int vocab = 0, max_pos = 0, hidden = 0, heads = 0, ffn = 0, n_layers = 0;
  float eps = 1e-12f;
  int precision = LK_F32;  // LK_F32 = split-bf16 operands (three MMAs per product), LK_BF16 = plain bf16 operands
  float *word = nullptr, *pos = nullptr, *type0 = nullptr, *emb_g = nullptr, *emb_b = nullptr;
  std::vector<BertLayer> layers;
  std::vector<void*> owned;  // every device allocation of the weights
  int* err_flag = nullptr;
  Buf ids, mask, x, qkv, tmp, planes_h, planes_f, out;  // planes_*: operand planes of [tokens, hidden] / [tokens, ffn]
};

namespace {

int upload(lk_bert* m, const void* src, size_t bytes, void** dst) {
  *dst = nullptr;
  if (!src) {
    set_error("lk_bert_create: a weight pointer is null");
    return LK_ERR_INVALID;
  }
  cudaError_t e = cudaMalloc(dst, bytes);
  if (e != cudaSuccess) return cuda_fail(e, "cudaMalloc(bert weights)", __FILE__, __LINE__);
  m->owned.push_back(*dst);
  LK_CUDA(cudaMemcpy(*dst, src, bytes, cudaMemcpyHostToDevice));
  return LK_OK;
}

template <typename T>
int upload_f(lk_bert* m, const float* src, size_t n, T** dst) {
  return upload(m, src, n * sizeof(float), reinterpret_cast<void**>(dst));
}

// nn.Linear weight [n_out, k_in] -> slab planes on the device
int upload_linear(lk_bert* m, const float* w, int n_out, int k_in, unsigned char** dst) {
  if (!w) {
    set_error("lk_bert_create: a weight pointer is null");
    return LK_ERR_INVALID;
  }
  std::vector<unsigned char> slabs;
  ae_umma_weight_slabs(w, n_out, k_in, kBlockRows, &slabs, false);
  return upload(m, slabs.data(), slabs.size(), reinterpret_cast<void**>(dst));
}

}  // namespace

extern "C" {

int lk_bert_destroy(lk_bert* m) {
  if (!m) return LK_OK;
  DeviceGuard guard(m->device);
  for (void* p : m->owned) cudaFree(p);
  if (m->err_flag) cudaFree(m->err_flag);
  Buf* bufs[] = {&m->ids, &m->mask, &m->x, &m->qkv, &m->tmp, &m->planes_h, &m->planes_f, &m->out};
  for (Buf* b : bufs) b->release();
  delete m;
  return LK_OK;
}

int lk_bert_create(lk_bert** out, int device, int vocab, int max_pos, int hidden, int heads, int ffn, int n_layers,
                   float ln_eps, const lk_bert_weights* w) {
  if (!out) return LK_ERR_INVALID;
  *out = nullptr;
  if (!w || !w->layers || vocab < 1 || max_pos < 1 || n_layers < 1 || n_layers > 64 || !(ln_eps >= 0.f)) {
    set_error("lk_bert_create: bad argument");
    return LK_ERR_INVALID;
  }
  if (!bert_shape_supported(hidden, heads, ffn) || !gemm_umma_supported(hidden, hidden) ||
      !gemm_umma_supported(ffn, hidden)) {
    set_error("lk_bert_create: hidden=%d heads=%d ffn=%d is outside what the kernels implement (head dimension 32, "
              "hidden and ffn multiples of 128, hidden <= 1024)", hidden, heads, ffn);
    return LK_ERR_UNSUPPORTED;
  }
  int sm = 0;
  int rc = check_device(device, &sm);
  if (rc != LK_OK) return rc;
  DeviceGuard guard(device);
  lk_bert* m = new (std::nothrow) lk_bert();
  if (!m) return LK_ERR_OOM;
  m->device = device;
  m->sm_count = sm;
  m->vocab = vocab;
  m->max_pos = max_pos;
  m->hidden = hidden;
  m->heads = heads;
  m->ffn = ffn;
  m->n_layers = n_layers;
  m->eps = ln_eps;
  const size_t h = (size_t)hidden, f = (size_t)ffn;
  auto fail = [&](int code) {
    lk_bert_destroy(m);
    return code;
  };
  if ((rc = upload_f(m, w->word_emb, (size_t)vocab * h, &m->word)) != LK_OK) return fail(rc);
  if ((rc = upload_f(m, w->pos_emb, (size_t)max_pos * h, &m->pos)) != LK_OK) return fail(rc);
  if ((rc = upload_f(m, w->type_emb, h, &m->type0)) != LK_OK) return fail(rc);
  if ((rc = upload_f(m, w->emb_ln_g, h, &m->emb_g)) != LK_OK) return fail(rc);
  if ((rc = upload_f(m, w->emb_ln_b, h, &m->emb_b)) != LK_OK) return fail(rc);
  m->layers.resize((size_t)n_layers);
  std::vector<float> wqkv(3 * h * h), bqkv(3 * h);
  for (int l = 0; l < n_layers; ++l) {
    const lk_bert_layer_weights& s = w->layers[l];
    BertLayer& d = m->layers[(size_t)l];
    if (!s.wq || !s.wk || !s.wv || !s.bq || !s.bk || !s.bv) {
      set_error("lk_bert_create: a weight pointer of layer %d is null", l);
      return fail(LK_ERR_INVALID);
    }
    // one [3 hidden, hidden] projection: Q | K | V
    memcpy(wqkv.data(), s.wq, h * h * sizeof(float));
    memcpy(wqkv.data() + h * h, s.wk, h * h * sizeof(float));
    memcpy(wqkv.data() + 2 * h * h, s.wv, h * h * sizeof(float));
    memcpy(bqkv.data(), s.bq, h * sizeof(float));
    memcpy(bqkv.data() + h, s.bk, h * sizeof(float));
    memcpy(bqkv.data() + 2 * h, s.bv, h * sizeof(float));
    if ((rc = upload_linear(m, wqkv.data(), 3 * hidden, hidden, &d.wqkv)) != LK_OK) return fail(rc);
    if ((rc = upload_f(m, bqkv.data(), 3 * h, &d.bqkv)) != LK_OK) return fail(rc);
    if ((rc = upload_linear(m, s.wo, hidden, hidden, &d.wo)) != LK_OK) return fail(rc);
    if ((rc = upload_f(m, s.bo, h, &d.bo)) != LK_OK) return fail(rc);
    if ((rc = upload_f(m, s.ln1_g, h, &d.ln1_g)) != LK_OK) return fail(rc);
    if ((rc = upload_f(m, s.ln1_b, h, &d.ln1_b)) != LK_OK) return fail(rc);
    if ((rc = upload_linear(m, s.w1, ffn, hidden, &d.w1)) != LK_OK) return fail(rc);
    if ((rc = upload_f(m, s.b1, f, &d.b1)) != LK_OK) return fail(rc);
    if ((rc = upload_linear(m, s.w2, hidden, ffn, &d.w2)) != LK_OK) return fail(rc);
    if ((rc = upload_f(m, s.b2, h, &d.b2)) != LK_OK) return fail(rc);
    if ((rc = upload_f(m, s.ln2_g, h, &d.ln2_g)) != LK_OK) return fail(rc);
    if ((rc = upload_f(m, s.ln2_b, h, &d.ln2_b)) != LK_OK) return fail(rc);
  }
  cudaError_t e = cudaMalloc((void**)&m->err_flag, sizeof(int));
  if (e == cudaSuccess) e = cudaMemset(m->err_flag, 0, sizeof(int));
  if (e != cudaSuccess) return fail(cuda_fail(e, "bert error flag", __FILE__, __LINE__));
  *out = m;
  return LK_OK;
}

int lk_bert_set_precision(lk_bert* m, int precision) {
  if (!m || (precision != LK_F32 && precision != LK_BF16)) return LK_ERR_INVALID;
  m->precision = precision;
  return LK_OK;
}

int lk_bert_encode(lk_bert* m, const int32_t* input_ids, const int32_t* attention_mask, int ids_mem, int64_t n_sent,
                   int seq_len, int normalize, float* out, int out_mem, void* stream) {
  if (!m || n_sent < 0 || seq_len < 1 || (ids_mem != LK_HOST && ids_mem != LK_DEVICE) ||
      (out_mem != LK_HOST && out_mem != LK_DEVICE) || (n_sent > 0 && (!input_ids || !attention_mask || !out))) {
    set_error("lk_bert_encode: bad argument");
    return LK_ERR_INVALID;
  }
  if (seq_len > m->max_pos) {
    set_error("lk_bert_encode: %d tokens per sentence exceed the %d position embeddings", seq_len, m->max_pos);
    return LK_ERR_INVALID;
  }
  if (n_sent == 0) return LK_OK;
  DeviceGuard guard(m->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int h = m->hidden, f = m->ffn, np = m->precision == LK_BF16 ? 1 : 2;
  int64_t per = kChunkTokens / seq_len;  // sentences per pass
  if (per < 1) per = 1;
  if (per > 65535) per = 65535;  // sentences are a grid dimension of the attention kernels
  if (per > n_sent) per = n_sent;
  const int64_t t_max = per * seq_len, t_pad = round_up64(t_max, kBlockRows);
  int rc;
  if ((rc = m->x.ensure((size_t)t_max * h * 4)) != LK_OK || (rc = m->qkv.ensure((size_t)t_max * 3 * h * 4)) != LK_OK ||
      (rc = m->tmp.ensure((size_t)t_max * h * 4)) != LK_OK ||
      (rc = m->planes_h.ensure((size_t)t_pad * h * 4)) != LK_OK ||  // 2 planes x 2 bytes per element
      (rc = m->planes_f.ensure((size_t)t_pad * f * 4)) != LK_OK)
    return rc;
  if (ids_mem == LK_HOST)
    if ((rc = m->ids.ensure((size_t)t_max * 4)) != LK_OK || (rc = m->mask.ensure((size_t)t_max * 4)) != LK_OK) return rc;
  if (out_mem == LK_HOST)
    if ((rc = m->out.ensure((size_t)per * h * 4)) != LK_OK) return rc;
  float* x = m->x.as<float>();
  float* tmp = m->tmp.as<float>();
  // Every activation a linear layer reads is written as operand planes by the kernel that produces
  // it (LayerNorm, attention, the GELU layer's epilogue): ph holds x / the context / the
  // post-attention state in turn, pf the feed-forward activations.
  unsigned char* ph = m->planes_h.as<unsigned char>();
  unsigned char* pf = m->planes_f.as<unsigned char>();
  auto linear = [&](const unsigned char* in, int k, const unsigned char* w, int n, const float* bias,
                    const float* residual, int act, float* y, unsigned char* y_planes, int64_t t) {
    return launch_gemm_umma(in, t, k, w, n, bias, residual, act, np, y, y_planes, m->err_flag, m->sm_count, st);
  };

  for (int64_t s0 = 0; s0 < n_sent; s0 += per) {
    const int64_t ns = n_sent - s0 < per ? n_sent - s0 : per, t = ns * seq_len;
    const int32_t* ids = input_ids + s0 * seq_len;
    const int32_t* mask = attention_mask + s0 * seq_len;
    if (ids_mem == LK_HOST) {
      LK_CUDA(cudaMemcpyAsync(m->ids.p, ids, (size_t)t * 4, cudaMemcpyHostToDevice, st));
      LK_CUDA(cudaMemcpyAsync(m->mask.p, mask, (size_t)t * 4, cudaMemcpyHostToDevice, st));
      ids = m->ids.as<int32_t>();
      mask = m->mask.as<int32_t>();
    }
    rc = launch_bert_embed_ln(ids, t, seq_len, m->vocab, h, m->word, m->pos, m->type0, m->emb_g, m->emb_b, m->eps, x, ph,
                              np, st);
    if (rc != LK_OK) return rc;
    for (const BertLayer& L : m->layers) {
      if ((rc = linear(ph, h, L.wqkv, 3 * h, L.bqkv, nullptr, 0, m->qkv.as<float>(), nullptr, t)) != LK_OK) return rc;
      if ((rc = launch_bert_attention(m->qkv.as<float>(), mask, ns, seq_len, h, m->heads, ph, np, st)) != LK_OK) return rc;
      if ((rc = linear(ph, h, L.wo, h, L.bo, x, 0, tmp, nullptr, t)) != LK_OK) return rc;
      if ((rc = launch_bert_layernorm(tmp, t, h, L.ln1_g, L.ln1_b, m->eps, ph, np, st)) != LK_OK) return rc;
      if ((rc = linear(ph, h, L.w1, f, L.b1, nullptr, 1, nullptr, pf, t)) != LK_OK) return rc;
      if ((rc = linear(pf, f, L.w2, h, L.b2, tmp, 0, x, nullptr, t)) != LK_OK) return rc;
      if ((rc = launch_bert_layernorm(x, t, h, L.ln2_g, L.ln2_b, m->eps, ph, np, st)) != LK_OK) return rc;
    }
    float* dst = out_mem == LK_HOST ? m->out.as<float>() : out + s0 * h;
    if ((rc = launch_bert_pool(x, mask, ns, seq_len, h, normalize, dst, st)) != LK_OK) return rc;
    if (out_mem == LK_HOST) {
      LK_CUDA(cudaMemcpyAsync(out + s0 * h, dst, (size_t)ns * h * 4, cudaMemcpyDeviceToHost, st));
      LK_CUDA(cudaStreamSynchronize(st));  // the staging buffers are reused by the next pass
    } else if (ids_mem == LK_HOST && s0 + per < n_sent) {
      LK_CUDA(cudaStreamSynchronize(st));
    }
  }
  if (out_mem == LK_HOST || ids_mem == LK_HOST) {
    int flag = 0;
    LK_CUDA(cudaMemcpyAsync(&flag, m->err_flag, sizeof(int), cudaMemcpyDeviceToHost, st));
    LK_CUDA(cudaStreamSynchronize(st));
    if (flag != 0) {
      cudaMemsetAsync(m->err_flag, 0, sizeof(int), st);
      set_error("linear-layer kernel pipeline timed out (barrier code %d); results are invalid", flag);
      return LK_ERR_CUDA;
    }
  }
  return LK_OK;
}

// One linear layer y = act(x W^T + b) (+ residual) through the same split / tcgen05 kernels, all
// operands fp32 on the HOST: the unit-test and bring-up entry of the encoder's GEMMs.
int lk_linear_forward(int device, const float* x, int64_t m, int k, const float* w, int n, const float* bias,
                      const float* residual, int act, int precision, float* y) {
  if (!x || !w || !y || m < 1 || act < 0 || act > 1 || (precision != LK_F32 && precision != LK_BF16)) {
    set_error("lk_linear_forward: bad argument");
    return LK_ERR_INVALID;
  }
  if (!gemm_umma_supported(n, k)) {
    set_error("lk_linear_forward: n=%d must be a multiple of 128 and k=%d a multiple of 64", n, k);
    return LK_ERR_UNSUPPORTED;
  }
  int sm = 0;
  int rc = check_device(device, &sm);
  if (rc != LK_OK) return rc;
  DeviceGuard guard(device);
  std::vector<unsigned char> slabs;
  ae_umma_weight_slabs(w, n, k, kBlockRows, &slabs, false);
  Buf dx, dw, db, dr, dy, dp, de;
  auto done = [&](int code) {
    Buf* bufs[] = {&dx, &dw, &db, &dr, &dy, &dp, &de};
    for (Buf* b : bufs) b->release();
    return code;
  };
  const size_t xb = (size_t)m * k * 4, yb = (size_t)m * n * 4;
  if ((rc = dx.ensure(xb)) != LK_OK || (rc = dw.ensure(slabs.size())) != LK_OK || (rc = db.ensure((size_t)n * 4)) != LK_OK ||
      (rc = dr.ensure(yb)) != LK_OK || (rc = dy.ensure(yb)) != LK_OK ||
      (rc = dp.ensure((size_t)round_up64(m, kBlockRows) * k * 4)) != LK_OK || (rc = de.ensure(sizeof(int))) != LK_OK)
    return done(rc);
  cudaError_t e = cudaMemcpy(dx.p, x, xb, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(dw.p, slabs.data(), slabs.size(), cudaMemcpyHostToDevice);
  if (e == cudaSuccess && bias) e = cudaMemcpy(db.p, bias, (size_t)n * 4, cudaMemcpyHostToDevice);
  if (e == cudaSuccess && residual) e = cudaMemcpy(dr.p, residual, yb, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemset(de.p, 0, sizeof(int));
  if (e != cudaSuccess) return done(cuda_fail(e, "lk_linear_forward upload", __FILE__, __LINE__));
  const int np = precision == LK_BF16 ? 1 : 2;
  if ((rc = launch_ae_split_rows(dx.as<float>(), m, k, np, dp.as<unsigned char>(), nullptr)) != LK_OK) return done(rc);
  rc = launch_gemm_umma(dp.as<unsigned char>(), m, k, dw.as<unsigned char>(), n, bias ? db.as<float>() : nullptr,
                        residual ? dr.as<float>() : nullptr, act, np, dy.as<float>(), nullptr, de.as<int>(), sm, nullptr);
  if (rc != LK_OK) return done(rc);
  int flag = 0;
  e = cudaMemcpy(y, dy.p, yb, cudaMemcpyDeviceToHost);
  if (e == cudaSuccess) e = cudaMemcpy(&flag, de.p, sizeof(int), cudaMemcpyDeviceToHost);
  if (e != cudaSuccess) return done(cuda_fail(e, "lk_linear_forward download", __FILE__, __LINE__));
  if (flag != 0) {
    set_error("linear-layer kernel pipeline timed out (barrier code %d); results are invalid", flag);
    return done(LK_ERR_CUDA);
  }
  return done(LK_OK);
}

int lk_bert_check(lk_bert* m) {
  if (!m) return LK_ERR_INVALID;
  DeviceGuard guard(m->device);
  int flag = 0;
  LK_CUDA(cudaMemcpy(&flag, m->err_flag, sizeof(int), cudaMemcpyDeviceToHost));  // synchronises
  if (flag != 0) {
    cudaMemset(m->err_flag, 0, sizeof(int));
    set_error("linear-layer kernel pipeline timed out (barrier code %d); results are invalid", flag);
    return LK_ERR_CUDA;
  }
  return LK_OK;
}

}  // extern "C"
