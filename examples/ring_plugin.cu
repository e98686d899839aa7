// examples/ring_plugin.cu -- a user likelihood plugin, start to finish.
//
// The stateline way (src/bin/demo-worker.cpp:37-45, src/app/workerwrapper.hpp:21):
//     double myNLL(uint jobType, const std::vector<double>& x) { ... }
//     WorkerWrapper w(myNLL, {0, nJobTypes}, "host:5555");  w.start();
// The B200 way: the same function body as a __device__ functor, compiled into a plugin .so:
//     nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -O3 -fmad=false -shared -Xcompiler -fPIC \
//          -I<repo>/include -o libring_plugin.so ring_plugin.cu
// and handed to slgpu_create(config_json, "libring_plugin.so", NULL, &payload, sizeof payload, &engine).
//
// Target: a 2-job factorisation of a "ring" posterior in d dimensions,
//   job 0: (|x| - radius)^2 / (2 width^2)        job 1: +inf outside the ball |x| <= cutoff, else 0
// (job 1 exercises the "+inf is a certain reject" rule, src/infer/chainarray.cpp:36-37).
#include "slgpu_device.cuh"

struct RingPayload {  // POD, copied to the device by the engine
  double radius, width, cutoff;
};

struct RingLik {
  const RingPayload* p;
  __device__ explicit RingLik(const void* payload) : p((const RingPayload*)payload) {}
  __device__ double operator()(uint32_t job, const double* x, uint32_t d) const {
    double r2 = 0.0;
    for (uint32_t i = 0; i < d; ++i) r2 += x[i] * x[i];
    const double r = sqrt(r2);
    if (job == 0) {
      const double t = (r - p->radius) / p->width;
      return 0.5 * (t * t);
    }
    return r <= p->cutoff ? 0.0 : __longlong_as_double(0x7ff0000000000000LL);
  }
};

// table symbol, functor, name, nJobTypes (0 = any), nDims (0 = any), payload bytes, dimensions to specialise
SLGPU_DEFINE_PLUGIN(slgpu_plugin_entry, RingLik, "ring", 2, 0, sizeof(RingPayload), 2, 3)
