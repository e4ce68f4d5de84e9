// Replays a fixed script of data-carrier operations (Tensor / TensorWrapper<T> / TensorMap, SURVEY.md 8a row a22 and the error rules of 8b)
// and prints what happens.  Built twice from this same source: against the reference's src/utils/tensor.h where it lies
// (oracle/Makefile, target ref_tensor -> oracle/_ref/tensor_ref: the checker) and against the shim's header
// (__graft_entry__.build() -> shim/_own_programs/tensor_shim: the product).  CPU objects only; nothing is ever destroyed (the reference's
// TensorWrapper destructor frees `data`).  Test infrastructure.
#include <cstdio>
#include <iostream>
#include <stdexcept>
#include <string>
#include <vector>
#include TENSOR_HEADER

// the reference's messages end in "Assertion fail: <file>:<line>": keep the text in front of it
static std::string essence(const std::exception &e) {
    std::string s = e.what();
    const size_t at = s.find("Assertion fail:");
    if (at != std::string::npos) s = s.substr(0, at);
    while (!s.empty() && (s.back() == ' ' || s.back() == '\n')) s.pop_back();
    return s;
}
#define SHOW(label, expr)                                                         \
    do {                                                                          \
        try {                                                                     \
            std::cout << label << " = " << (expr) << "\n";                        \
        } catch (const std::exception &e) {                                       \
            std::cout << label << " THROWS std::exception: " << essence(e) << "\n"; \
        }                                                                         \
    } while (0)

int main() {
    float *fbuf = new float[24];
    int *ibuf = new int[4];
    for (int i = 0; i < 24; ++i) fbuf[i] = 0.5f * i;
    for (int i = 0; i < 4; ++i) ibuf[i] = 10 + i;
    auto *a = new TensorWrapper<float>(Device::CPU, DataType::FP32, {2, 3, 4}, fbuf);
    auto *step = new TensorWrapper<int>(Device::CPU, DataType::INT32, {1}, ibuf);
    auto *gpu = new TensorWrapper<float>(Device::GPU, DataType::FP32, {4, 6}, fbuf);  // declared on the device: host reads must be refused
    auto *no_data = new TensorWrapper<float>(Device::CPU, DataType::FP32, {2, 2}, nullptr);
    auto *no_shape = new TensorWrapper<float>(Device::CPU, DataType::FP32, {}, fbuf);
    auto *plain = new Tensor(Device::CPU, DataType::INT32, {5, 7});

    SHOW("size(a)", a->size());
    SHOW("size(step)", step->size());
    SHOW("size(no_data)", no_data->size());
    SHOW("size(no_shape)", no_shape->size());
    SHOW("size(plain)", plain->size());
    SHOW("a.getVal(5)", a->getVal(5));
    SHOW("step.getVal()", step->getVal());
    SHOW("gpu.getVal()", gpu->getVal());
    SHOW("a.getPtrByOffset(3) - a.getPtr()", (long)(a->getPtrByOffset(3) - a->getPtr()));
    SHOW("plain.toString()", plain->toString());
    SHOW("a.deviceString()", a->deviceString());
    SHOW("gpu.deviceString()", gpu->deviceString());
    SHOW("getTensorType<float>", (int)getTensorType<float>());
    SHOW("getTensorType<const int>", (int)getTensorType<const int>());
    SHOW("getTensorType<bool>", (int)getTensorType<bool>());
    SHOW("getTensorType<double>", (int)getTensorType<double>());
    SHOW("dtype mismatch ctor", (new TensorWrapper<float>(Device::CPU, DataType::INT32, {1}, fbuf))->size());
    SHOW("wrap<float>(a) same object", (long)(static_cast<Tensor *>(a)->wrap<float>() == a));

    TensorMap m{{"decoder_input", a}, {"step", step}};
    SHOW("m.size()", m.size());
    SHOW("m.isExist(step)", m.isExist("step"));
    SHOW("m.isExist(finished)", m.isExist("finished"));
    SHOW("m.at(step)->size()", m.at("step")->size());
    SHOW("m[decoder_input]->size()", m["decoder_input"]->size());
    {   // a missing key names itself (the key list that follows depends on hash order: cut at "(keys")
        try {
            m.at("finished");
            std::cout << "m.at(finished) = found\n";
        } catch (const std::exception &e) {
            std::string s = essence(e);
            const size_t at = s.find("(keys");
            std::cout << "m.at(finished) THROWS std::exception: " << (at == std::string::npos ? s : s.substr(0, at)) << "\n";
        }
    }
    m.insert("step", gpu);  // insert overwrites
    SHOW("m.at(step)->size() after overwrite", m.at("step")->size());
    m.insert({"layer_id", plain});
    SHOW("m.size() after inserts", m.size());
    SHOW("invalid tensor in the initialiser list", (TensorMap{{"attention_input", a}, {"broken", no_data}}).size());
    SHOW("isValid(no_shape)", m.isValid(no_shape));
    SHOW("isValid(a)", m.isValid(a));
    m.insert("unchecked", no_data);  // insert() itself does not validate
    SHOW("m.size() after inserting an invalid tensor", m.size());
    m.insert(std::pair<std::string, Tensor *>("layer_id", a));  // a pair does not replace an existing entry
    SHOW("m.at(layer_id)->size() after pair insert", m.at("layer_id")->size());
    SHOW("vec2str({1,2,3})", vec2str(std::vector<int>{1, 2, 3}));
    SHOW("vec2str({})", vec2str(std::vector<int>{}));
    SHOW("vec2str({x})", vec2str(std::vector<std::string>{"x"}));
    {
        const int arr[3] = {4, 5, 6};
        const int *ap = arr;
        SHOW("arr2str(3)", arr2str(ap, 3));
    }
    SHOW("fmtstr", fmtstr("%s-%d-%.2f", "k", 7, 1.5));
    std::unordered_map<std::string, Tensor *> src{{"x", a}, {"empty", no_data}};
    SHOW("from unordered_map (invalid entries skipped)", (new TensorMap(src))->size());
    return 0;
}
