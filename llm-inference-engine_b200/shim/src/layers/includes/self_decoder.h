// Drop-in for the reference's src/layers/includes/self_decoder.h:15-84 + src/layers/self_decoder.cpp:24-122: the decode step
// over all layers.
//   inputs  {"decoder_input" [bs, hidden], "step" (CPU int), "finished" [bs] bool, "layer_id" (CPU int)} (+ "output_norm_weight",
//            passed by the reference's callers and unused)
//   outputs {"decoder_output" [bs, hidden], "all_k_cache", "all_v_cache" [L, bs, Hkv, S, d]}
// Two execution paths with the same results:
//   * composition (any weights): RMSNorm -> LlamaSelfAttentionLayer -> fused add-bias-residual-RMSNorm -> LlamaFFNLayer ->
//     add-residual per layer, exactly the reference's launcher sequence, but stream-ordered: no per-op malloc / free / sync;
//   * fused engine (every layer packed with LlamaLayerWeight::packForB200()): b200_decoder_step -- five PDL-chained kernels
//     per layer (include/b200llm.h), CUDA-graph capturable.
#pragma once

#include <vector>
#include "../../kernels/includes/decoder_self_attention.cuh"
#include "../../kernels/includes/add_residual_and_rmsnorm.cuh"
#include "../../kernels/includes/rmsnorm.cuh"
#include "../../kernels/includes/add_residual.cuh"
#include "../../layers/includes/self_attention.h"
#include "../../layers/includes/ffn.h"
#include "../../weights/includes/llama_weights.h"
#include "../../utils/tensor.h"

template <typename T> class LlamaSelfDecoder {
private:
    int head_num;
    int kv_head_num;
    int head_size;
    int intermediate_size;
    int num_layer;
    int hidden_units;
    float rmsnorm_eps;
    LlamaAttentionStaticParams attn_params;  // copied: the reference keeps a pointer to the caller's (often temporary) struct

    cudaStream_t stream;
    CublasWrapper *cublas_wrapper;
    BaseAllocator *allocator;

    TensorWrapper<T> *decoder_residual = nullptr;
    LlamaSelfAttentionLayer<T> *self_attention = nullptr;
    LlamaFFNLayer<T> *ffn = nullptr;
    DataType data_type;
    b200shim::Workspace workspace;
    cudaStream_t active_stream = nullptr;

    // fused engine state
    b200_decoder_t *engine = nullptr;
    b200shim::Workspace engine_scratch;
    int engine_batch = 0, engine_seq = 0;
    std::vector<const void *> engine_weights;  // identity of the packed buffers the engine was built on

    bool allPacked(std::vector<LlamaLayerWeight<T> *> *lw) const {
        if ((int)lw->size() < num_layer) return false;
        for (int l = 0; l < num_layer; ++l) {
            LlamaLayerWeight<T> *w = lw->at(l);
            const BaseWeight<T> *ws[4] = {&w->self_attention_weight.qkv, &w->self_attention_weight.output, &w->ffn_weight.gate_and_up, &w->ffn_weight.down};
            for (const BaseWeight<T> *b : ws)
                if (!b->packed || b->packed_from != b->data || b->packed_type != ws[0]->packed_type) return false;
        }
        return true;
    }
    static int wformat(WeightType t) { return t == WeightType::FP8_W ? B200_W_FP8E4M3 : (t == WeightType::INT4_W ? B200_W_INT4 : B200_W_DENSE); }

    void buildEngine(std::vector<LlamaLayerWeight<T> *> *lw, int max_batch, int max_seq) {
        std::vector<const void *> ident;
        for (int l = 0; l < num_layer; ++l) ident.push_back(lw->at(l)->self_attention_weight.qkv.packed);
        if (engine && engine_batch == max_batch && engine_seq == max_seq && ident == engine_weights) return;
        if (engine) b200_decoder_destroy(engine);
        b200_decoder_config_t c = {};
        c.hidden = hidden_units, c.head_num = head_num, c.kv_head_num = kv_head_num, c.head_size = head_size;
        c.inter_size = intermediate_size, c.num_layers = num_layer, c.max_seq_len = max_seq, c.max_batch = max_batch;
        c.dtype = b200DType<T>();
        c.w_format = wformat(lw->at(0)->self_attention_weight.qkv.packed_type);
        c.group = lw->at(0)->self_attention_weight.qkv.group_size;
        c.rmsnorm_eps = rmsnorm_eps;
        c.rotary_dim = attn_params.rotary_embedding_dim < head_size ? attn_params.rotary_embedding_dim : head_size;
        c.rotary_base = attn_params.rotary_embedding_base;
        c.tp_world = 1, c.tp_rank = 0;
        engine = b200_decoder_create(&c);
        LLM_CHECK_WITH_INFO(engine != nullptr, std::string("b200_decoder_create: ") + b200_last_error_string());
        for (int l = 0; l < num_layer; ++l) {
            LlamaLayerWeight<T> *w = lw->at(l);
            auto lin = [](const BaseWeight<T> &b) { return b200_linear_weight_t{b.packed, b.packed_scales, b.packed_zeros}; };
            b200_layer_weights_t lwt = {};
            lwt.attn_norm_gamma = w->attention_norm_weight.gamma;
            lwt.qkv = lin(w->self_attention_weight.qkv), lwt.qkv_bias = w->self_attention_weight.qkv.bias;
            lwt.o = lin(w->self_attention_weight.output), lwt.o_bias = w->self_attention_weight.output.bias;
            lwt.ffn_norm_gamma = w->ffn_norm_weight.gamma;
            lwt.gate_up = lin(w->ffn_weight.gate_and_up), lwt.down = lin(w->ffn_weight.down);
            B200_CALL(b200_decoder_set_layer(engine, l, &lwt));
        }
        const size_t bytes = b200_decoder_scratch_bytes(engine);
        engine_scratch.reserve(bytes);
        B200_CALL(b200_decoder_set_scratch(engine, engine_scratch.template take<char>(bytes), bytes));
        engine_batch = max_batch, engine_seq = max_seq, engine_weights = ident;
    }

public:
    LlamaSelfDecoder(const int &head_num, const int &kv_head_num, const int &head_size, const int &intermediate_size, const int &num_layer,
                     const LlamaAttentionStaticParams &attn_params, const float &rmsnorm_eps, const cudaStream_t &stream,
                     CublasWrapper *const &cublas_wrapper, BaseAllocator *const &allocator)
        : head_num(head_num), kv_head_num(kv_head_num), head_size(head_size), intermediate_size(intermediate_size), num_layer(num_layer),
          hidden_units(head_num * head_size), rmsnorm_eps(rmsnorm_eps), attn_params(attn_params), stream(stream), cublas_wrapper(cublas_wrapper),
          allocator(allocator), data_type(getTensorType<T>()), workspace(allocator), engine_scratch(allocator) {
        self_attention = new LlamaSelfAttentionLayer<T>(head_num, kv_head_num, head_size, &this->attn_params, stream, cublas_wrapper, allocator);
        ffn = new LlamaFFNLayer<T>(head_num, head_size, intermediate_size, stream, cublas_wrapper, allocator);
    }
    ~LlamaSelfDecoder() {
        freeBuf();
        delete self_attention;
        delete ffn;
        if (engine) b200_decoder_destroy(engine);
    }
    LlamaSelfDecoder(const LlamaSelfDecoder &) = delete;
    LlamaSelfDecoder &operator=(const LlamaSelfDecoder &) = delete;

    // New: run every launcher of forward() on `s` (the reference stores its constructor's stream and never uses it).
    void setStream(cudaStream_t s) {
        active_stream = s;
        self_attention->setStream(s);
        ffn->setStream(s);
    }

    void allocateMemory(LlamaAttentionDynamicParams *dynamic_params) {
        const size_t n = (size_t)dynamic_params->batch_size * hidden_units;
        if (decoder_residual && decoder_residual->shape[0] == dynamic_params->batch_size) return;
        workspace.reserve(b200shim::Workspace::padded(n, sizeof(T)));
        delete decoder_residual;
        decoder_residual = new TensorWrapper<T>(Device::GPU, data_type, {dynamic_params->batch_size, hidden_units}, workspace.template take<T>(n));
    }
    void freeBuf() {
        delete decoder_residual;
        decoder_residual = nullptr;
        workspace.release();
    }

    void forward(TensorMap *input_tensors, std::vector<LlamaLayerWeight<T> *> *layer_weights, TensorMap *output_tensors,
                 LlamaAttentionDynamicParams *dynamic_params) {
        b200shim::StreamScope scope(active_stream);
        allocateMemory(dynamic_params);
        Tensor *decoder_input = input_tensors->at("decoder_input");
        Tensor *step = input_tensors->at("step");
        Tensor *finished = input_tensors->at("finished");
        Tensor *decoder_output = output_tensors->at("decoder_output");
        Tensor *all_k_cache = output_tensors->at("all_k_cache");
        Tensor *all_v_cache = output_tensors->at("all_v_cache");
        Tensor *layer_id = input_tensors->at("layer_id");
        LLM_CHECK_WITH_INFO(decoder_input->wrap<T>()->data != nullptr, "The data pointer of tensor inserted into TensorMap is nullptr!");
        LLM_CHECK_WITH_INFO(step->wrap<int>()->data != nullptr, "The data pointer of tensor inserted into TensorMap is nullptr!");
        LLM_CHECK_WITH_INFO(finished->wrap<bool>()->data != nullptr, "The data pointer of tensor inserted into TensorMap is nullptr!");

        const int bs = dynamic_params->batch_size;
        if (allPacked(layer_weights) && all_k_cache->shape.size() == 5) {
            // ---- fused engine: decoder_output <- decoder_input, then the residual stream is updated in place
            buildEngine(layer_weights, all_k_cache->shape[1], all_k_cache->shape[3]);
            T *out = decoder_output->wrap<T>()->data;
            if (out != decoder_input->wrap<T>()->data)
                CHECK(cudaMemcpyAsync(out, decoder_input->wrap<T>()->data, sizeof(T) * (size_t)bs * hidden_units, cudaMemcpyDeviceToDevice,
                                      b200GetStream()));
            B200_CALL(b200_decoder_step(engine, out, all_k_cache->wrap<T>()->data, all_v_cache->wrap<T>()->data, bs, step->wrap<int>()->getVal(), 0,
                                        num_layer, b200GetStream()));
            return;
        }

        // ---- composition: the reference's launcher sequence (self_decoder.cpp:69-119)
        TensorMap self_attention_inputs{{"attention_input", decoder_input}, {"layer_id", layer_id}, {"step", step}, {"finished", finished}};
        TensorMap self_attention_outputs{{"attention_output", decoder_output}, {"all_k_cache", all_k_cache}, {"all_v_cache", all_v_cache}};
        std::vector<int> ids(num_layer);
        std::vector<TensorWrapper<int> *> id_tensors;
        for (int l = 0; l < num_layer; ++l) {
            if (l > 0) {
                ids[l] = l;
                id_tensors.push_back(new TensorWrapper<int>(Device::CPU, getTensorType<int>(), std::vector<int>{1}, &ids[l]));
                self_attention_inputs.insert("layer_id", id_tensors.back());  // (key, value): overwrites -- a pair would keep layer 0 (tensor.h:241-243)
            }
            Tensor *x = self_attention_inputs.at("attention_input");
            launchRMSNorm(x->wrap<T>(), decoder_residual, &layer_weights->at(l)->attention_norm_weight, rmsnorm_eps);
            self_attention->forward(&self_attention_inputs, &self_attention_outputs, &layer_weights->at(l)->self_attention_weight, dynamic_params);
            launchFusedAddBiasResidualAndRMSNorm(decoder_residual, decoder_output->wrap<T>(), &layer_weights->at(l)->self_attention_weight.output,
                                                 layer_weights->at(l)->ffn_norm_weight.gamma, rmsnorm_eps);
            TensorMap ffn_inputs{{"ffn_input", decoder_output}};
            TensorMap ffn_outputs{{"ffn_output", decoder_output}};
            ffn->forward(&ffn_inputs, &ffn_outputs, &layer_weights->at(l)->ffn_weight, dynamic_params);
            launchAddResidual(decoder_residual, decoder_output->wrap<T>(), true);
            self_attention_inputs.insert("attention_input", decoder_output);
        }
        for (TensorWrapper<int> *t : id_tensors) delete t;  // views of ids[]: nothing else is freed (TensorMap does not own)
    }
};
