/*
 * ORACLE support (test infrastructure, NOT product code): a minimal stand-in for the parts of the
 * TensorFlow 1.4 C++ custom-op API that the reference's op sources use, so that
 * /root/reference/lmbspecialops/src/{warp2d,median3x3downsample,scaleinvariantgradient,leakyrelu,depthtoflow}.cc
 * compile UNMODIFIED, from where they lie, into oracle/_ref/libref_ops.so (TensorFlow itself cannot be installed here).
 *
 * What runs in _ref is the reference's own OpKernel::Compute() body: index math, loops, branches, border and NaN
 * handling.  What this header supplies is only plumbing: attribute lookup, tensors as (shape, buffer), a registry that
 * REGISTER_KERNEL_BUILDER fills and oracle/ref_harness.cc instantiates kernels from.  REGISTER_OP(...) chains (attribute
 * declarations, shape functions, doc strings) compile but are never executed.
 *
 * API surface mirrored (tensorflow/core/framework/{op.h,op_kernel.h,shape_inference.h} of TF 1.4.0, the version the
 * reference pins, Dockerfile:14): OpKernel, OpKernelConstruction::GetAttr, OpKernelContext::{input,allocate_output},
 * Tensor::{shape,flat<T>}, TensorShape::{dims,dim_size,set_dim,AddDim,num_elements}, Status, errors::InvalidArgument,
 * OP_REQUIRES(_OK), REGISTER_OP, REGISTER_KERNEL_BUILDER, shape_inference::InferenceContext (signatures only).
 */
#ifndef ORACLE_TF_STUB_H
#define ORACLE_TF_STUB_H

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <functional>
#include <initializer_list>
#include <limits>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "eigen_stub.h"   // TensorFlow's framework headers pull in Eigen; warp2d.cc relies on that

namespace tensorflow {

typedef int64_t int64;

class Status {
 public:
  Status() : ok_(true) {}
  explicit Status(const std::string& m) : ok_(false), msg_(m) {}
  static Status OK() { return Status(); }
  bool ok() const { return ok_; }
  const std::string& error_message() const { return msg_; }
 private:
  bool ok_;
  std::string msg_;
};

namespace errors {
template <class... Args>
inline Status InvalidArgument(const char* m, Args...) { return Status(std::string(m)); }
}  // namespace errors

#define TF_RETURN_IF_ERROR(expr)                 \
  do {                                           \
    ::tensorflow::Status _st = (expr);           \
    if (!_st.ok()) return _st;                   \
  } while (0)

// ---- shape inference: signatures only (the lambdas passed to SetShapeFn compile, nothing calls them) ----------------
namespace shape_inference {
struct ShapeHandle {};
struct DimensionHandle {};
struct DimensionOrConstant {
  DimensionOrConstant(DimensionHandle) {}
  DimensionOrConstant(int64) {}
  DimensionOrConstant(int) {}
};
class InferenceContext {
 public:
  ShapeHandle input(int) { return ShapeHandle(); }
  void set_output(int, ShapeHandle) {}
  Status WithRank(ShapeHandle, int64, ShapeHandle*) { return Status::OK(); }
  Status WithRankAtLeast(ShapeHandle, int64, ShapeHandle*) { return Status::OK(); }
  Status WithRankAtMost(ShapeHandle, int64, ShapeHandle*) { return Status::OK(); }
  Status WithValue(DimensionHandle, int64, DimensionHandle*) { return Status::OK(); }
  bool RankKnown(ShapeHandle) { return false; }
  bool ValueKnown(DimensionOrConstant) { return false; }
  int64 Value(DimensionOrConstant) { return 0; }
  int Rank(ShapeHandle) { return 0; }
  DimensionHandle Dim(ShapeHandle, int64) { return DimensionHandle(); }
  Status Subshape(ShapeHandle, int64, ShapeHandle*) { return Status::OK(); }
  Status Subshape(ShapeHandle, int64, int64, ShapeHandle*) { return Status::OK(); }
  Status Merge(ShapeHandle, ShapeHandle, ShapeHandle*) { return Status::OK(); }
  Status Merge(DimensionHandle, DimensionHandle, DimensionHandle*) { return Status::OK(); }
  Status Concatenate(ShapeHandle, ShapeHandle, ShapeHandle*) { return Status::OK(); }
  Status ReplaceDim(ShapeHandle, int64, DimensionHandle, ShapeHandle*) { return Status::OK(); }
  Status Multiply(DimensionHandle, DimensionOrConstant, DimensionHandle*) { return Status::OK(); }
  Status Add(DimensionHandle, DimensionOrConstant, DimensionHandle*) { return Status::OK(); }
  Status Divide(DimensionHandle, DimensionOrConstant, bool, DimensionHandle*) { return Status::OK(); }
  DimensionHandle MakeDim(DimensionOrConstant) { return DimensionHandle(); }
  DimensionHandle UnknownDim() { return DimensionHandle(); }
  ShapeHandle MakeShape(std::initializer_list<DimensionOrConstant>) { return ShapeHandle(); }
  ShapeHandle MakeShape(const std::vector<DimensionHandle>&) { return ShapeHandle(); }
  ShapeHandle UnknownShape() { return ShapeHandle(); }
  ShapeHandle Scalar() { return ShapeHandle(); }
  ShapeHandle Vector(DimensionOrConstant) { return ShapeHandle(); }
  ShapeHandle Matrix(DimensionOrConstant, DimensionOrConstant) { return ShapeHandle(); }
  template <class T>
  Status GetAttr(const std::string&, T*) { return Status::OK(); }
};
}  // namespace shape_inference

// ---- REGISTER_OP("X").Attr(..).Input(..).Output(..).SetShapeFn(..).Doc(..): swallowed ----------------------------------
class OpDefBuilderStub {
 public:
  explicit OpDefBuilderStub(const char*) {}
  OpDefBuilderStub& Attr(const char*) { return *this; }
  OpDefBuilderStub& Input(const char*) { return *this; }
  OpDefBuilderStub& Output(const char*) { return *this; }
  OpDefBuilderStub& Doc(const char*) { return *this; }
  OpDefBuilderStub& SetIsStateful() { return *this; }
  template <class F>
  OpDefBuilderStub& SetShapeFn(F) { return *this; }
};
#define ORACLE_TF_CAT2(a, b) a##b
#define ORACLE_TF_CAT(a, b) ORACLE_TF_CAT2(a, b)
#define REGISTER_OP(name) \
  static ::tensorflow::OpDefBuilderStub ORACLE_TF_CAT(oracle_opdef_, __COUNTER__) __attribute__((unused)) = ::tensorflow::OpDefBuilderStub(name)

// ---- tensors -------------------------------------------------------------------------------------------------------
class TensorShape {
 public:
  TensorShape() {}
  explicit TensorShape(const std::vector<int64>& d) : d_(d) {}
  int dims() const { return (int)d_.size(); }
  int64 dim_size(int i) const { return d_[i]; }
  void set_dim(int i, int64 v) { d_[i] = v; }
  void AddDim(int64 v) { d_.push_back(v); }
  int64 num_elements() const { int64 n = 1; for (int64 v : d_) n *= v; return n; }
  const std::vector<int64>& vec() const { return d_; }
 private:
  std::vector<int64> d_;
};

template <class T>
struct FlatView {
  T* p;
  int64 n;
  T* data() const { return p; }
  int64 size() const { return n; }
  T& operator()(int64 i) const { return p[i]; }
};

class Tensor {
 public:
  Tensor() : buf_(nullptr), bytes_(0) {}
  Tensor(const TensorShape& s, void* borrowed) : shape_(s), buf_(borrowed), bytes_(0) {}
  Tensor(const TensorShape& s, size_t elem_bytes) : shape_(s), bytes_(s.num_elements() * elem_bytes) {
    own_.reset(new unsigned char[bytes_ ? bytes_ : 1]);
    buf_ = own_.get();
  }
  const TensorShape& shape() const { return shape_; }
  int dims() const { return shape_.dims(); }
  int64 dim_size(int i) const { return shape_.dim_size(i); }
  int64 NumElements() const { return shape_.num_elements(); }
  template <class T> FlatView<const T> flat() const { return FlatView<const T>{static_cast<const T*>(buf_), shape_.num_elements()}; }
  template <class T> FlatView<T> flat() { return FlatView<T>{static_cast<T*>(buf_), shape_.num_elements()}; }
  void* raw() const { return buf_; }
 private:
  TensorShape shape_;
  void* buf_;
  size_t bytes_;
  std::shared_ptr<unsigned char> own_;
};

class PersistentTensor {};   // only members of that type exist in the CPU kernels, never used there

// ---- kernels -------------------------------------------------------------------------------------------------------
struct AttrValue {
  bool has_b = false, has_f = false, has_i = false, has_s = false, has_li = false, has_lf = false;
  bool b = false;
  float f = 0.f;
  int64 i = 0;
  std::string s;
  std::vector<int> li;
  std::vector<float> lf;
};

class OpKernelConstruction {
 public:
  std::map<std::string, AttrValue> attrs;
  Status status;
  void SetStatus(const Status& s) { if (status.ok()) status = s; }
  void CtxFailureWithWarning(const Status& s) { SetStatus(s); }
  Status GetAttr(const std::string& n, bool* v) const { return get(n, [&](const AttrValue& a) { *v = a.b; return a.has_b; }); }
  Status GetAttr(const std::string& n, float* v) const { return get(n, [&](const AttrValue& a) { *v = a.f; return a.has_f; }); }
  Status GetAttr(const std::string& n, int* v) const { return get(n, [&](const AttrValue& a) { *v = (int)a.i; return a.has_i; }); }
  Status GetAttr(const std::string& n, int64* v) const { return get(n, [&](const AttrValue& a) { *v = a.i; return a.has_i; }); }
  Status GetAttr(const std::string& n, std::string* v) const { return get(n, [&](const AttrValue& a) { *v = a.s; return a.has_s; }); }
  Status GetAttr(const std::string& n, std::vector<int>* v) const { return get(n, [&](const AttrValue& a) { *v = a.li; return a.has_li; }); }
  Status GetAttr(const std::string& n, std::vector<float>* v) const { return get(n, [&](const AttrValue& a) { *v = a.lf; return a.has_lf; }); }
 private:
  template <class F>
  Status get(const std::string& n, F f) const {
    auto it = attrs.find(n);
    if (it == attrs.end() || !f(it->second)) return Status("attribute '" + n + "' missing or of another type");
    return Status::OK();
  }
};

class OpKernelContext {
 public:
  std::vector<Tensor> inputs;
  std::vector<Tensor> outputs;
  size_t elem_bytes = 4;
  Status status;
  const Tensor& input(int i) { return inputs[i]; }
  int num_inputs() const { return (int)inputs.size(); }
  Status allocate_output(int i, const TensorShape& s, Tensor** out) {
    if ((int)outputs.size() <= i) outputs.resize(i + 1);
    outputs[i] = Tensor(s, elem_bytes);
    *out = &outputs[i];
    return Status::OK();
  }
  void SetStatus(const Status& s) { if (status.ok()) status = s; }
  void CtxFailureWithWarning(const Status& s) { SetStatus(s); }
};

class OpKernel {
 public:
  explicit OpKernel(OpKernelConstruction*) {}
  virtual ~OpKernel() {}
  virtual void Compute(OpKernelContext* context) = 0;
};

#define OP_REQUIRES_OK(CTX, ...)                          \
  do {                                                    \
    ::tensorflow::Status _s(__VA_ARGS__);                 \
    if (!_s.ok()) { (CTX)->SetStatus(_s); return; }       \
  } while (0)
#define OP_REQUIRES(CTX, EXP, STATUS)                     \
  do {                                                    \
    if (!(EXP)) { (CTX)->SetStatus(STATUS); return; }     \
  } while (0)

static const char* const DEVICE_CPU = "CPU";
static const char* const DEVICE_GPU = "GPU";

struct KernelDefStub {
  std::string op, device, dtype;
};
class Name {
 public:
  explicit Name(const char* op) { def_.op = op; }
  Name& Device(const char* d) { def_.device = d; return *this; }
  template <class T> Name& TypeConstraint(const char*);
  Name& HostMemory(const char*) { return *this; }
  const KernelDefStub& def() const { return def_; }
 private:
  KernelDefStub def_;
};
template <> inline Name& Name::TypeConstraint<float>(const char*) { def_.dtype = "float"; return *this; }
template <> inline Name& Name::TypeConstraint<double>(const char*) { def_.dtype = "double"; return *this; }

typedef OpKernel* (*KernelFactory)(OpKernelConstruction*);
struct KernelRegistry {
  static std::map<std::string, KernelFactory>& table() {
    static std::map<std::string, KernelFactory> t;
    return t;
  }
};
struct KernelRegistrar {
  KernelRegistrar(const Name& n, KernelFactory f) { KernelRegistry::table()[n.def().op + "/" + n.def().device + "/" + n.def().dtype] = f; }
};
#define REGISTER_KERNEL_BUILDER(kernel_builder, ...)                                                         \
  static ::tensorflow::KernelRegistrar ORACLE_TF_CAT(oracle_kernel_, __COUNTER__)(                           \
      ::tensorflow::kernel_builder,                                                                            \
      [](::tensorflow::OpKernelConstruction* c) -> ::tensorflow::OpKernel* { return new __VA_ARGS__(c); })

}  // namespace tensorflow
#endif
