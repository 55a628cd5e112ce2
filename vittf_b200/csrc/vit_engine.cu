// ViT K-feature engine: the device-side replacement of the hooked hub-model forward in
// /root/reference/infer.py:130-210 (compute_qkv).  Only what reaches the hook on
// blocks[-1].attn.qkv is evaluated: L-1 full blocks, then norm1 and the K third of the last
// block's qkv projection for the patch tokens.  The full 3D-wide qkv tensor the
// reference copies to the host per batch (infer.py:134) never exists.
#include <new>
#include <utility>
#include <vector>

#include "common.cuh"

struct vittf_vit {
    vittf_vit_config cfg;
    vittf_block_weights* blocks;  // host array, device pointers inside
    const float* patch_w;
    const float* patch_b;
    int max_batch, max_tokens;
    // optional per-kernel timing (bench.py roofline): event pairs around attention (0) / GEMM (1) launches
    bool timing = false;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> pending[2];
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> pool;
};

namespace {
inline int64_t align256(int64_t v) { return (v + 255) / 256 * 256; }
inline int tok_pad_of(int tokens) { return (tokens + 127) / 128 * 128; }

struct Workspace {
    float* x;             // residual stream fp32 (M, D)
    __nv_bfloat16* xn;    // LayerNorm output (M, D)
    __nv_bfloat16* qk;    // (M, 2D)
    __nv_bfloat16* vt;    // (B*heads*64, tok_pad)
    __nv_bfloat16* att;   // attention output (M, D)
    __nv_bfloat16* hid;   // MLP hidden (M, 4D)
    float* stats;         // LayerNorm fold: (m_pad, VITTF_LN_SLOTS) partial (sum, sum of squares) of the rows of x
    int64_t m_pad;
    void* attn_ws;        // per-CTA flags of the two-pass attention
    int64_t attn_ws_bytes;
    int64_t vt_bytes;
    int64_t total;
};

Workspace carve(const vittf_vit_config& c, int batch, int tokens, uint8_t* base) {
    Workspace w;
    const int64_t M = static_cast<int64_t>(batch) * tokens;
    const int D = c.embed_dim;
    int64_t off = 0;
    auto take = [&](int64_t bytes) { uint8_t* p = base ? base + off : nullptr; off += align256(bytes); return p; };
    w.m_pad = (M + 255) / 256 * 256;          // the row-tiled stream of the LayerNorm-folded path holds whole 256-row pair tiles
    w.x = reinterpret_cast<float*>(take(w.m_pad * D * 4));
    w.xn = reinterpret_cast<__nv_bfloat16*>(take(M * D * 2));
    w.stats = reinterpret_cast<float*>(take(w.m_pad * VITTF_LN_SLOTS * 8));
    w.qk = reinterpret_cast<__nv_bfloat16*>(take(M * 2 * D * 2));
    w.vt_bytes = static_cast<int64_t>(batch) * D * tok_pad_of(tokens) * 2;
    w.vt = reinterpret_cast<__nv_bfloat16*>(take(w.vt_bytes));
    w.att = reinterpret_cast<__nv_bfloat16*>(take(M * D * 2));
    w.hid = reinterpret_cast<__nv_bfloat16*>(take(M * c.mlp_hidden * 2));
    w.attn_ws_bytes = vittf_attention_workspace_bytes(batch, tokens, c.num_heads);
    w.attn_ws = take(w.attn_ws_bytes);
    w.total = off;
    return w;
}

struct ScopedTimer {
    vittf_vit* v;
    int kind;
    cudaStream_t s;
    std::pair<cudaEvent_t, cudaEvent_t> ev{nullptr, nullptr};
    ScopedTimer(vittf_vit* v_, int kind_, cudaStream_t s_) : v(v_), kind(kind_), s(s_) {
        if (!v->timing) return;
        if (!v->pool.empty()) { ev = v->pool.back(); v->pool.pop_back(); }
        else { cudaEventCreate(&ev.first); cudaEventCreate(&ev.second); }
        cudaEventRecord(ev.first, s);
    }
    ~ScopedTimer() {
        if (!ev.first) return;
        cudaEventRecord(ev.second, s);
        v->pending[kind].push_back(ev);
    }
};
}  // namespace

extern "C" int vittf_vit_timing_enable(vittf_vit* v, int enable) {
    VITTF_REQUIRE(v, "vittf_vit_timing_enable: null engine");
    v->timing = enable != 0;
    return VITTF_OK;
}

extern "C" int vittf_vit_timing_read(vittf_vit* v, double* ms_by_kind2, int64_t* launches_by_kind2) {
    VITTF_REQUIRE(v && ms_by_kind2 && launches_by_kind2, "vittf_vit_timing_read: null pointer");
    for (int k = 0; k < 2; ++k) {
        double total = 0.0;
        for (auto& e : v->pending[k]) {
            VITTF_CHECK_CUDA(cudaEventSynchronize(e.second));
            float ms = 0.0f;
            VITTF_CHECK_CUDA(cudaEventElapsedTime(&ms, e.first, e.second));
            total += ms;
            v->pool.push_back(e);
        }
        ms_by_kind2[k] = total;
        launches_by_kind2[k] = static_cast<int64_t>(v->pending[k].size());
        v->pending[k].clear();
    }
    return VITTF_OK;
}

extern "C" int vittf_vit_create(vittf_vit** out, const vittf_vit_config* cfg, const vittf_block_weights* blocks_host,
                                const float* patch_w, const float* patch_b, int max_batch, int max_tokens) {
    VITTF_REQUIRE(out && cfg && blocks_host && patch_w && patch_b, "vittf_vit_create: null pointer");
    VITTF_REQUIRE(cfg->depth >= 1, "vittf_vit_create: depth must be >= 1");
    VITTF_REQUIRE(cfg->embed_dim == cfg->num_heads * 64, "vittf_vit_create: head dim must be 64 (D=%d heads=%d)",
                  cfg->embed_dim, cfg->num_heads);
    VITTF_REQUIRE(cfg->embed_dim % 128 == 0 && cfg->mlp_hidden % 128 == 0, "vittf_vit_create: D and MLP width must be multiples of 128");
    VITTF_REQUIRE(max_batch > 0 && max_tokens > 1, "vittf_vit_create: bad capacity");
    vittf_vit* v = new (std::nothrow) vittf_vit;
    VITTF_REQUIRE(v, "vittf_vit_create: out of host memory");
    v->cfg = *cfg;
    v->blocks = new (std::nothrow) vittf_block_weights[cfg->depth];
    if (!v->blocks) { delete v; VITTF_REQUIRE(false, "vittf_vit_create: out of host memory"); }
    for (int i = 0; i < cfg->depth; ++i) v->blocks[i] = blocks_host[i];
    for (int i = 0; i < cfg->depth; ++i) {
        const bool fold = blocks_host[i].qkv_colsum != nullptr;
        if (fold != (blocks_host[0].qkv_colsum != nullptr) || fold != (blocks_host[i].fc1_colsum != nullptr)) {
            delete[] v->blocks;
            delete v;
            VITTF_REQUIRE(false, "vittf_vit_create: qkv_colsum / fc1_colsum must be set for every block or for none (block %d)", i);
        }
    }
    if (blocks_host[0].qkv_colsum != nullptr && (cfg->mlp_hidden < 2 * cfg->embed_dim || vittf_gemm_ln_slots(cfg->embed_dim) > VITTF_LN_SLOTS)) {
        delete[] v->blocks;
        delete v;
        VITTF_REQUIRE(false, "vittf_vit_create: the LayerNorm-folded path needs mlp_hidden >= 2 D (token staging) and D = %d within %d "
                      "partial-sum slots", cfg->embed_dim, VITTF_LN_SLOTS);
    }
    v->patch_w = patch_w;
    v->patch_b = patch_b;
    v->max_batch = max_batch;
    v->max_tokens = max_tokens;
    *out = v;
    return VITTF_OK;
}

extern "C" void vittf_vit_destroy(vittf_vit* v) {
    if (!v) return;
    for (int k = 0; k < 2; ++k)
        for (auto& e : v->pending[k]) v->pool.push_back(e);
    for (auto& e : v->pool) { cudaEventDestroy(e.first); cudaEventDestroy(e.second); }
    delete[] v->blocks;
    delete v;
}

extern "C" int64_t vittf_vit_workspace_bytes(const vittf_vit* v, int batch, int tokens) {
    if (!v || batch <= 0 || tokens <= 1) return -1;
    return carve(v->cfg, batch, tokens, nullptr).total;
}

extern "C" int vittf_vit_k_features(vittf_vit* v, const void* vol, int vol_dtype, int X, int Y, int Z, int axis, int s0,
                                    int s1, int im0, int im1, const float* minmax2, const float* pos_embed,
                                    void* out_k_f16, void* workspace, int64_t workspace_bytes, void* stream) {
    VITTF_REQUIRE(v && vol && minmax2 && pos_embed && out_k_f16 && workspace, "vittf_vit_k_features: null pointer");
    const vittf_vit_config& c = v->cfg;
    const int B = s1 - s0;
    VITTF_REQUIRE(B > 0 && B <= v->max_batch, "vittf_vit_k_features: batch %d outside (0,%d]", B, v->max_batch);
    VITTF_REQUIRE(im0 % c.patch == 0 && im1 % c.patch == 0, "vittf_vit_k_features: image size not a multiple of the patch size");
    const int tokens = 1 + (im0 / c.patch) * (im1 / c.patch);
    VITTF_REQUIRE(tokens <= v->max_tokens, "vittf_vit_k_features: %d tokens exceed the engine capacity %d", tokens, v->max_tokens);
    Workspace w = carve(c, B, tokens, static_cast<uint8_t*>(workspace));
    if (w.total > workspace_bytes) {
        vittf_set_error("vittf_vit_k_features: workspace of %lld B is smaller than the %lld B required", (long long)workspace_bytes, (long long)w.total);
        return VITTF_ERR_NOMEM;
    }
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int D = c.embed_dim;
    const int M = B * tokens;
    const int tok_pad = tok_pad_of(tokens);
    // padding columns [tokens, tok_pad) of V^T are read (times P = 0) by the last key block: they must be finite.  Only they
    // are cleared (B * D rows of tok_pad - tokens elements); the QKV epilogue writes every other element.
    if (tok_pad > tokens)
        VITTF_CHECK_CUDA(cudaMemset2DAsync(w.vt + tokens, static_cast<size_t>(tok_pad) * 2, 0, static_cast<size_t>(tok_pad - tokens) * 2,
                                           static_cast<size_t>(B) * D, s));
    const bool fold = v->blocks[0].qkv_colsum != nullptr;
    if (fold) {
        // LayerNorm folded into the GEMMs (gemm.cu): x = row-tiled fp32 stream, xn = its raw bf16 copy, stats = row sums.
        // The patch embedding writes row-major tokens into the (still unused) MLP hidden buffer; one pass re-tiles them.
        float* x_rm = reinterpret_cast<float*>(w.hid);
        VITTF_CHECK(vittf_patch_embed(vol, vol_dtype, X, Y, Z, axis, s0, s1, im0, im1, c.patch, D, minmax2, v->patch_w,
                                      v->patch_b, pos_embed, x_rm, stream));
        VITTF_CHECK(vittf_ln_prepare(x_rm, w.x, w.xn, w.stats, M, w.m_pad, D, stream));
        vittf_ln_fold consume{};
        consume.stats = w.stats; consume.eps = 1e-6f; consume.m_pad = w.m_pad;
        vittf_ln_fold produce{};
        produce.xt = w.x; produce.stats_out = w.stats; produce.m_pad = w.m_pad;
        for (int l = 0; l + 1 < c.depth; ++l) {
            const vittf_block_weights& bw = v->blocks[l];
            consume.colsum = bw.qkv_colsum;
            { ScopedTimer t(v, 1, s);
              VITTF_CHECK(vittf_gemm_bf16_ln(w.xn, bw.qkv_w, bw.qkv_b, w.qk, w.vt, M, 3 * D, D, VITTF_EPI_QKV_SPLIT, tokens, tok_pad, &consume, stream)); }
            { ScopedTimer t(v, 0, s);
              VITTF_CHECK(vittf_attention_prescaled(w.qk, w.vt, w.att, B, tokens, tok_pad, c.num_heads, w.attn_ws, w.attn_ws_bytes, stream)); }
            { ScopedTimer t(v, 1, s);
              VITTF_CHECK(vittf_gemm_bf16_ln(w.att, bw.proj_w, bw.proj_b, w.xn, nullptr, M, D, D, VITTF_EPI_BIAS_RESID_LN, tokens, tok_pad, &produce, stream)); }
            consume.colsum = bw.fc1_colsum;
            { ScopedTimer t(v, 1, s);
              VITTF_CHECK(vittf_gemm_bf16_ln(w.xn, bw.fc1_w, bw.fc1_b, w.hid, nullptr, M, c.mlp_hidden, D, VITTF_EPI_BIAS_GELU_BF16, tokens, tok_pad, &consume, stream)); }
            { ScopedTimer t(v, 1, s);
              VITTF_CHECK(vittf_gemm_bf16_ln(w.hid, bw.fc2_w, bw.fc2_b, w.xn, nullptr, M, D, c.mlp_hidden, VITTF_EPI_BIAS_RESID_LN, tokens, tok_pad, &produce, stream)); }
        }
        // last block: norm1 + K rows [D, 2D) of attn.qkv only
        const vittf_block_weights& last = v->blocks[c.depth - 1];
        consume.colsum = last.qkv_colsum + D;
        const __nv_bfloat16* wk = static_cast<const __nv_bfloat16*>(last.qkv_w) + static_cast<size_t>(D) * D;
        VITTF_CHECK(vittf_gemm_bf16_ln(w.xn, wk, last.qkv_b + D, out_k_f16, nullptr, M, D, D, VITTF_EPI_KFEAT_F16, tokens, tok_pad, &consume, stream));
        return VITTF_OK;
    }
    VITTF_CHECK(vittf_patch_embed(vol, vol_dtype, X, Y, Z, axis, s0, s1, im0, im1, c.patch, D, minmax2, v->patch_w,
                                  v->patch_b, pos_embed, w.x, stream));
    for (int l = 0; l + 1 < c.depth; ++l) {
        const vittf_block_weights& bw = v->blocks[l];
        VITTF_CHECK(vittf_layernorm(w.x, bw.ln1_w, bw.ln1_b, w.xn, M, D, stream));
        { ScopedTimer t(v, 1, s);
          VITTF_CHECK(vittf_gemm_bf16(w.xn, bw.qkv_w, bw.qkv_b, w.qk, w.vt, M, 3 * D, D, VITTF_EPI_QKV_SPLIT, tokens, tok_pad, stream)); }
        { ScopedTimer t(v, 0, s);
          VITTF_CHECK(vittf_attention_prescaled(w.qk, w.vt, w.att, B, tokens, tok_pad, c.num_heads, w.attn_ws, w.attn_ws_bytes, stream)); }
        { ScopedTimer t(v, 1, s);
          VITTF_CHECK(vittf_gemm_bf16(w.att, bw.proj_w, bw.proj_b, w.x, nullptr, M, D, D, VITTF_EPI_BIAS_RESID_F32, tokens, tok_pad, stream)); }
        VITTF_CHECK(vittf_layernorm(w.x, bw.ln2_w, bw.ln2_b, w.xn, M, D, stream));
        { ScopedTimer t(v, 1, s);
          VITTF_CHECK(vittf_gemm_bf16(w.xn, bw.fc1_w, bw.fc1_b, w.hid, nullptr, M, c.mlp_hidden, D, VITTF_EPI_BIAS_GELU_BF16, tokens, tok_pad, stream)); }
        { ScopedTimer t(v, 1, s);
          VITTF_CHECK(vittf_gemm_bf16(w.hid, bw.fc2_w, bw.fc2_b, w.x, nullptr, M, D, c.mlp_hidden, VITTF_EPI_BIAS_RESID_F32, tokens, tok_pad, stream)); }
    }
    // last block: norm1 + K rows [D, 2D) of attn.qkv only
    const vittf_block_weights& last = v->blocks[c.depth - 1];
    VITTF_CHECK(vittf_layernorm(w.x, last.ln1_w, last.ln1_b, w.xn, M, D, stream));
    const __nv_bfloat16* wk = static_cast<const __nv_bfloat16*>(last.qkv_w) + static_cast<size_t>(D) * D;
    VITTF_CHECK(vittf_gemm_bf16(w.xn, wk, last.qkv_b + D, out_k_f16, nullptr, M, D, D, VITTF_EPI_KFEAT_F16, tokens, tok_pad, stream));
    return VITTF_OK;
}
