/*
 * ORACLE support (test infrastructure, NOT product code): C entry point that instantiates a kernel the reference's own
 * REGISTER_KERNEL_BUILDER registered (through oracle/ref_stub/tf_stub.h) and runs its Compute() on host buffers.
 * Linked with the UNMODIFIED reference sources into oracle/_ref/libref_ops.so by oracle/Makefile (target `ref`).
 */
#include <cstdio>
#include <cstring>
#include <sstream>
#include <string>

#include "tf_stub.h"

using namespace tensorflow;

static bool parse_attrs(const char* spec, std::map<std::string, AttrValue>& out, std::string& err) {
  // "name:type=value;..." with type in {b, f, i, s, li, lf}; list values separated by ','
  std::stringstream ss(spec ? spec : "");
  std::string item;
  while (std::getline(ss, item, ';')) {
    if (item.empty()) continue;
    const size_t c = item.find(':'), e = item.find('=');
    if (c == std::string::npos || e == std::string::npos || e < c) { err = "bad attribute spec: " + item; return false; }
    const std::string name = item.substr(0, c), type = item.substr(c + 1, e - c - 1), val = item.substr(e + 1);
    AttrValue a;
    if (type == "b") { a.has_b = true; a.b = (val == "1" || val == "true"); }
    else if (type == "f") { a.has_f = true; a.f = std::strtof(val.c_str(), nullptr); }
    else if (type == "i") { a.has_i = true; a.i = std::strtoll(val.c_str(), nullptr, 10); }
    else if (type == "s") { a.has_s = true; a.s = val; }
    else if (type == "li" || type == "lf") {
      std::stringstream ls(val);
      std::string v;
      while (std::getline(ls, v, ',')) {
        if (v.empty()) continue;
        if (type == "li") a.li.push_back((int)std::strtol(v.c_str(), nullptr, 10)); else a.lf.push_back(std::strtof(v.c_str(), nullptr));
      }
      if (type == "li") a.has_li = true; else a.has_lf = true;
    } else { err = "bad attribute type: " + item; return false; }
    out[name] = a;
  }
  return true;
}

extern "C" {

/* Runs CPU kernel `op` ("Warp2d", "DepthToFlow", ...) for float (is_double = 0) or double.
 * shapes_flat holds the dims of all inputs back to back (ranks[i] each).  The output is copied to `out` (capacity in
 * elements); its shape goes to out_shape / out_rank.  Returns 0, or -1 with a message in err. */
int ref_run(const char* op, int is_double, const char* attr_spec, int ninputs, const void* const* data, const int64_t* shapes_flat,
            const int* ranks, void* out, int64_t out_capacity, int64_t* out_shape, int* out_rank, char* err, int errlen) {
  auto fail = [&](const std::string& m) { if (err && errlen > 0) snprintf(err, errlen, "%s", m.c_str()); return -1; };
  const std::string key = std::string(op) + "/CPU/" + (is_double ? "double" : "float");
  auto& table = KernelRegistry::table();
  auto it = table.find(key);
  if (it == table.end()) return fail("no such kernel registered by the reference sources: " + key);
  OpKernelConstruction cons;
  std::string perr;
  if (!parse_attrs(attr_spec, cons.attrs, perr)) return fail(perr);
  std::unique_ptr<OpKernel> kernel(it->second(&cons));
  if (!cons.status.ok()) return fail("construction failed: " + cons.status.error_message());
  OpKernelContext ctx;
  ctx.elem_bytes = is_double ? 8 : 4;
  const int64_t* sp = shapes_flat;
  for (int i = 0; i < ninputs; ++i) {
    std::vector<int64> dims(sp, sp + ranks[i]);
    sp += ranks[i];
    ctx.inputs.emplace_back(TensorShape(dims), const_cast<void*>(data[i]));
  }
  kernel->Compute(&ctx);
  if (!ctx.status.ok()) return fail("Compute failed: " + ctx.status.error_message());
  if (ctx.outputs.empty()) return fail("kernel produced no output");
  const Tensor& o = ctx.outputs[0];
  const int64_t n = o.shape().num_elements();
  if (n > out_capacity) return fail("output buffer too small");
  memcpy(out, o.raw(), (size_t)n * ctx.elem_bytes);
  *out_rank = o.shape().dims();
  for (int i = 0; i < o.shape().dims(); ++i) out_shape[i] = o.shape().dim_size(i);
  return 0;
}

/* number of kernels registered; names are written ';'-separated */
int ref_list(char* buf, int buflen) {
  std::string s;
  for (auto& kv : KernelRegistry::table()) s += kv.first + ";";
  snprintf(buf, buflen, "%s", s.c_str());
  return (int)KernelRegistry::table().size();
}

}  // extern "C"
