// oracle/ref_harness.cpp -- TEST INFRASTRUCTURE ONLY.
//
// A thin extern "C" driver over the UNMODIFIED reference's public C++ API, compiled
// against the headers where they lie under /root/reference (oracle/Makefile) and
// linked to oracle/_ref/libnmslib_ref.so.  It exists because the as-shipped C ABI
//   * pins efSearch=200 on every query       (nmslib_c.cpp:330, :986)
//   * builds HNSW on empty data / rebuilds it (lib.zig:629, nmslib_c.cpp:1682-1704)
//   * is a serial loop                        (nmslib_c.cpp:1015-1023)
// so the efSearch sweep, the "same graph" export (Hnsw::SaveIndex, hnsw.cc:748-806)
// and the OpenMP-over-queries CPU baseline (CPU-B) all need direct
// calls to Index<dist_t>::Search.  Only tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline / --impl reference legs may load this library.

#include <omp.h>

#include <cstdint>
#include <cstring>
#include <memory>
#include <sstream>
#include <string>
#include <vector>

#include "init.h"
#include "index.h"
#include "knnquery.h"
#include "knnqueue.h"
#include "methodfactory.h"
#include "object.h"
#include "params.h"
#include "space.h"
#include "space/space_l2sqr_sift.h"
#include "space/space_vector.h"
#include "spacefactory.h"

using namespace similarity;

namespace {

std::vector<std::string> split_csv(const char* csv) {
  std::vector<std::string> out;
  if (!csv) return out;
  std::stringstream ss(csv);
  std::string tok;
  while (std::getline(ss, tok, ','))
    if (!tok.empty()) out.push_back(tok);
  return out;
}

struct HarnessBase {
  virtual ~HarnessBase() {}
  virtual int add_f32(const float*, size_t, size_t, const int32_t*) { return -1; }
  virtual int add_u8(const uint8_t*, size_t, const int32_t*) { return -1; }
  virtual int build(const char* params) = 0;
  virtual int set_qparams(const char* params) = 0;
  virtual int knn(const void* q, size_t nq, size_t dim, size_t k, int32_t* ids, float* dists,
                  int32_t* counts, int threads) = 0;
  virtual int save(const char* path) = 0;
  virtual int load(const char* path) = 0;
  virtual size_t size() const = 0;
  std::string err;
};

template <typename dist_t>
struct Harness : HarnessBase {
  std::string space_name, method_name;
  std::unique_ptr<Space<dist_t>> space;
  std::unique_ptr<Index<dist_t>> index;
  ObjectVector data;
  bool is_u8 = false;

  ~Harness() override {
    index.reset();
    for (auto* o : data) delete o;
  }
  size_t size() const override { return data.size(); }

  Object* make_obj(const void* p, size_t dim, int32_t id) const {
    if (is_u8) {
      auto* sp = dynamic_cast<const SpaceL2SqrSift*>(space.get());
      const uint8_t* u = static_cast<const uint8_t*>(p);
      std::vector<uint8_t> v(u, u + dim);
      return sp->CreateObjFromUint8Vect(id, -1, v);
    } else {
      // same route as nmslib_c.cpp:235-243 (float payload behind a 16-byte header)
      const float* f = static_cast<const float*>(p);
      return new Object(id, -1, dim * sizeof(float), f);
    }
  }

  int add_f32(const float* d, size_t n, size_t dim, const int32_t* ids) override {
    if (is_u8) return -1;
    for (size_t i = 0; i < n; ++i)
      data.push_back(make_obj(d + i * dim, dim, ids ? ids[i] : (int32_t)(data.size())));
    return 0;
  }
  int add_u8(const uint8_t* d, size_t n, const int32_t* ids) override {
    if (!is_u8) return -1;
    for (size_t i = 0; i < n; ++i)
      data.push_back(make_obj(d + i * 128, 128, ids ? ids[i] : (int32_t)(data.size())));
    return 0;
  }
  int build(const char* params) override {
    try {
      index.reset(MethodFactoryRegistry<dist_t>::Instance().CreateMethod(
          false, method_name, space_name, *space, data));
      index->CreateIndex(AnyParams(split_csv(params)));
      return 0;
    } catch (const std::exception& e) {
      err = e.what();
      return -2;
    }
  }
  int set_qparams(const char* params) override {
    try {
      index->SetQueryTimeParams(AnyParams(split_csv(params)));
      return 0;
    } catch (const std::exception& e) {
      err = e.what();
      return -2;
    }
  }
  int knn(const void* q, size_t nq, size_t dim, size_t k, int32_t* ids, float* dists,
          int32_t* counts, int threads) override {
    if (!index) return -3;
    const size_t esz = is_u8 ? 1 : sizeof(float);
    int failed = 0;
    if (threads < 1) threads = 1;
#pragma omp parallel for schedule(dynamic, 16) num_threads(threads)
    for (long long i = 0; i < (long long)nq; ++i) {
      try {
        std::unique_ptr<Object> qo(make_obj((const char*)q + (size_t)i * dim * esz, dim, 0));
        KNNQuery<dist_t> knn(*space, qo.get(), (unsigned)k);
        index->Search(&knn);
        // same extraction order as nmslib_c.cpp:313-327: pop (descending), reverse
        std::unique_ptr<KNNQueue<dist_t>> res(knn.Result()->Clone());
        size_t found = res->Size();
        counts[i] = (int32_t)found;
        for (size_t j = found; j-- > 0;) {
          dists[i * k + j] = (float)res->TopDistance();
          ids[i * k + j] = (int32_t)res->TopObject()->id();
          res->Pop();
        }
      } catch (...) {
#pragma omp atomic write
        failed = 1;
      }
    }
    return failed ? -4 : 0;
  }
  int save(const char* path) override {
    try {
      index->SaveIndex(path);
      return 0;
    } catch (const std::exception& e) {
      err = e.what();
      return -2;
    }
  }
  // Index::LoadIndex on an empty data set, the way nmslib_load_index does it (nmslib_c.cpp:1428-1435): the
  // optimized HNSW stream carries the vectors itself.  Used to let the REFERENCE search a graph this library built.
  int load(const char* path) override {
    try {
      index.reset(MethodFactoryRegistry<dist_t>::Instance().CreateMethod(
          false, method_name, space_name, *space, data));
      index->LoadIndex(path);
      return 0;
    } catch (const std::exception& e) {
      err = e.what();
      return -2;
    }
  }
};

}  // namespace

extern "C" {

// is_int_dist: 0 -> Space<float> (dense float vectors), 1 -> Space<int> (l2sqr_sift uint8)
void* refh_open(const char* space, const char* method, int is_int_dist) {
  initLibrary(0, LIB_LOGNONE, nullptr);
  try {
    if (is_int_dist) {
      auto* h = new Harness<int>();
      h->space_name = space;
      h->method_name = method;
      h->is_u8 = true;
      h->space.reset(SpaceFactoryRegistry<int>::Instance().CreateSpace(space, AnyParams()));
      return static_cast<HarnessBase*>(h);
    }
    auto* h = new Harness<float>();
    h->space_name = space;
    h->method_name = method;
    h->space.reset(SpaceFactoryRegistry<float>::Instance().CreateSpace(space, AnyParams()));
    return static_cast<HarnessBase*>(h);
  } catch (...) {
    return nullptr;
  }
}
void refh_close(void* h) { delete static_cast<HarnessBase*>(h); }
int refh_add_f32(void* h, const float* d, size_t n, size_t dim, const int32_t* ids) {
  return static_cast<HarnessBase*>(h)->add_f32(d, n, dim, ids);
}
int refh_add_u8(void* h, const uint8_t* d, size_t n, const int32_t* ids) {
  return static_cast<HarnessBase*>(h)->add_u8(d, n, ids);
}
int refh_build(void* h, const char* params_csv) {
  return static_cast<HarnessBase*>(h)->build(params_csv);
}
int refh_set_query_params(void* h, const char* params_csv) {
  return static_cast<HarnessBase*>(h)->set_qparams(params_csv);
}
int refh_knn_batch(void* h, const void* q, size_t nq, size_t dim, size_t k, int32_t* ids,
                   float* dists, int32_t* counts, int threads) {
  return static_cast<HarnessBase*>(h)->knn(q, nq, dim, k, ids, dists, counts, threads);
}
int refh_save(void* h, const char* path) { return static_cast<HarnessBase*>(h)->save(path); }
int refh_load(void* h, const char* path) { return static_cast<HarnessBase*>(h)->load(path); }
size_t refh_size(void* h) { return static_cast<HarnessBase*>(h)->size(); }
const char* refh_last_error(void* h) { return static_cast<HarnessBase*>(h)->err.c_str(); }
int refh_max_threads(void) { return omp_get_max_threads(); }

}  // extern "C"
