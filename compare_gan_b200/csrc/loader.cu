// Host-side input pipeline: the tf.data chain of ImageDatasetV2.train_input_fn
// (reference datasets.py:261-291) — repeat -> shuffle(buffer, seed) -> batch(drop_remainder) -> prefetch — as one
// producer thread that fills a ring of page-locked batch buffers while the GPU runs the previous cycle.  No device work
// happens here; the consumer (ModularGAN.set_inputs) issues the host->device copies and releases the slots afterwards.
#include <cuda_runtime.h>

#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

#include "../../include/cgan_b200.h"

struct cgan_loader {
  const uint8_t* src_u8 = nullptr;
  const float* src_f32 = nullptr;
  const int32_t* src_labels = nullptr;
  int64_t n = 0;
  int64_t elems = 0;                 // h*w*c
  int batch = 0, ring = 0;
  std::vector<float*> images;        // ring slots
  std::vector<int32_t*> labels;
  bool pinned = false;
  float u8_to_unit[256];             // v / 255.0f, a true division as in TF (a multiply by 1/255 is 1 ulp off for some v)

  // tf.data shuffle: a buffer of element indices; the source is the infinite repeat() of 0..n-1
  std::vector<int64_t> shuffle;
  int64_t next_source = 0;
  uint64_t rng_state = 0;

  // ring protocol: slots [tail, head) are filled or outstanding; produced counts fills, consumed counts next() calls,
  // released counts slots handed back.  produced - released <= ring.
  std::mutex mu;
  std::condition_variable cv_producer, cv_consumer;
  int64_t produced = 0, consumed = 0, released = 0;
  bool stop = false;
  std::thread worker;
  char err[256] = {0};
};

namespace {

inline uint64_t splitmix64(uint64_t& s) {
  uint64_t z = (s += 0x9E3779B97F4A7C15ull);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

// uniform integer in [0, bound) without modulo bias (rejection on the top of the range)
inline uint64_t uniform_below(uint64_t& s, uint64_t bound) {
  const uint64_t limit = UINT64_MAX - UINT64_MAX % bound;
  uint64_t r;
  do { r = splitmix64(s); } while (r >= limit);
  return r % bound;
}

inline int64_t next_element(cgan_loader* L) {
  if (L->shuffle.empty()) {              // no shuffling: the plain repeat() stream
    int64_t e = L->next_source;
    L->next_source = (L->next_source + 1) % L->n;
    return e;
  }
  const uint64_t slot = uniform_below(L->rng_state, L->shuffle.size());
  const int64_t e = L->shuffle[slot];
  L->shuffle[slot] = L->next_source;     // replaced by the next input element
  L->next_source = (L->next_source + 1) % L->n;
  return e;
}

void fill_slot(cgan_loader* L, int slot) {
  float* img = L->images[slot];
  int32_t* lab = L->labels[slot];
  for (int b = 0; b < L->batch; ++b) {
    const int64_t e = next_element(L);
    float* dst = img + (int64_t)b * L->elems;
    if (L->src_u8) {
      const uint8_t* src = L->src_u8 + e * L->elems;
      for (int64_t i = 0; i < L->elems; ++i) dst[i] = L->u8_to_unit[src[i]];     // _parse_fn: tf.cast(image, float32) / 255.0
    } else {
      memcpy(dst, L->src_f32 + e * L->elems, (size_t)L->elems * sizeof(float));
    }
    lab[b] = L->src_labels ? L->src_labels[e] : 0;
  }
}

void producer(cgan_loader* L) {
  for (;;) {
    int slot;
    {
      std::unique_lock<std::mutex> lk(L->mu);
      L->cv_producer.wait(lk, [&] { return L->stop || L->produced - L->released < L->ring; });
      if (L->stop) return;
      slot = (int)(L->produced % L->ring);
    }
    fill_slot(L, slot);                  // outside the lock: the slot is neither filled nor outstanding
    {
      std::lock_guard<std::mutex> lk(L->mu);
      ++L->produced;
    }
    L->cv_consumer.notify_one();
  }
}

int lfail(cgan_loader* L, int code, const char* msg) {
  if (L) snprintf(L->err, sizeof(L->err), "%s", msg);
  return code;
}

}  // namespace

extern "C" {

int cgan_loader_create(cgan_loader** out, const void* images, int src_dtype, const int32_t* labels, int64_t n, int h, int w,
                       int c, int batch, int shuffle_buffer, uint64_t seed, int ring) {
  if (!out || !images || n < 1 || h < 1 || w < 1 || c < 1 || batch < 1 || ring < 2 || (src_dtype != 0 && src_dtype != 1))
    return CGAN_ERR_ARG;
  cgan_loader* L = new cgan_loader();
  if (src_dtype == 0) L->src_u8 = static_cast<const uint8_t*>(images);
  else L->src_f32 = static_cast<const float*>(images);
  L->src_labels = labels;
  L->n = n;
  L->elems = (int64_t)h * w * c;
  L->batch = batch;
  L->ring = ring;
  L->rng_state = seed;
  for (int v = 0; v < 256; ++v) L->u8_to_unit[v] = (float)v / 255.0f;
  if (shuffle_buffer > 1) {              // tf.data fills the buffer with the first `buffer` elements of the stream
    L->shuffle.resize((size_t)shuffle_buffer);
    for (int i = 0; i < shuffle_buffer; ++i) {
      L->shuffle[i] = L->next_source;
      L->next_source = (L->next_source + 1) % n;
    }
  }
  int ndev = 0;
  L->pinned = cudaGetDeviceCount(&ndev) == cudaSuccess && ndev > 0;
  if (!L->pinned) cudaGetLastError();    // clear the "no device" error: plain host memory is fine for a host pipeline
  const size_t ib = (size_t)batch * L->elems * sizeof(float), lb = (size_t)batch * sizeof(int32_t);
  for (int s = 0; s < ring; ++s) {
    void *pi = nullptr, *pl = nullptr;
    if (L->pinned) {
      if (cudaHostAlloc(&pi, ib, cudaHostAllocPortable) != cudaSuccess || cudaHostAlloc(&pl, lb, cudaHostAllocPortable) != cudaSuccess) {
        cgan_loader_destroy(L);
        return CGAN_ERR_CUDA;
      }
    } else {
      pi = aligned_alloc(64, (ib + 63) / 64 * 64);
      pl = aligned_alloc(64, (lb + 63) / 64 * 64);
      if (!pi || !pl) { cgan_loader_destroy(L); return CGAN_ERR_WORKSPACE; }
    }
    L->images.push_back(static_cast<float*>(pi));
    L->labels.push_back(static_cast<int32_t*>(pl));
  }
  L->worker = std::thread(producer, L);
  *out = L;
  return CGAN_OK;
}

int cgan_loader_next(cgan_loader* L, const float** images, const int32_t** labels) {
  if (!L || !images || !labels) return CGAN_ERR_ARG;
  std::unique_lock<std::mutex> lk(L->mu);
  if (L->consumed - L->released >= L->ring)
    return lfail(L, CGAN_ERR_ARG, "cgan_loader_next: every ring slot is outstanding; call cgan_loader_release first");
  L->cv_consumer.wait(lk, [&] { return L->produced > L->consumed; });
  const int slot = (int)(L->consumed % L->ring);
  ++L->consumed;
  *images = L->images[slot];
  *labels = L->labels[slot];
  return CGAN_OK;
}

int cgan_loader_release(cgan_loader* L, int count) {
  if (!L || count < 0) return CGAN_ERR_ARG;
  {
    std::lock_guard<std::mutex> lk(L->mu);
    if (L->released + count > L->consumed) return lfail(L, CGAN_ERR_ARG, "cgan_loader_release: more slots than outstanding");
    L->released += count;
  }
  L->cv_producer.notify_one();
  return CGAN_OK;
}

int cgan_loader_destroy(cgan_loader* L) {
  if (!L) return CGAN_ERR_ARG;
  {
    std::lock_guard<std::mutex> lk(L->mu);
    L->stop = true;
  }
  L->cv_producer.notify_all();
  if (L->worker.joinable()) L->worker.join();
  for (float* p : L->images) { if (L->pinned) cudaFreeHost(p); else free(p); }
  for (int32_t* p : L->labels) { if (L->pinned) cudaFreeHost(p); else free(p); }
  delete L;
  return CGAN_OK;
}

const char* cgan_loader_last_error(cgan_loader* L) { return L ? L->err : "null loader"; }

}  // extern "C"
