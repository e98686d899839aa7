/*
 * oracle/ref_shim/zmq.hpp -- declarations-only stand-in for cppzmq's zmq.hpp, found BEFORE the reference's
 * external/cppzmq on the include path of oracle/Makefile.ref (libzmq's C header zmq.h is not in this image).
 * TEST INFRASTRUCTURE ONLY.  src/infer/sampler.hpp pulls in comms/requester.hpp -> comms/socket.hpp, which
 * only needs the two types below to declare its members; no ZeroMQ call is compiled: the harness
 * (oracle/ref_harness.cpp) defines comms::Requester as an in-process FIFO instead of compiling src/comms.
 */
#ifndef SLREF_ZMQ_SHIM
#define SLREF_ZMQ_SHIM
#define ZMQ_DEALER 5
namespace zmq {
class context_t {
 public:
  explicit context_t(int = 1) {}
};
class socket_t {
 public:
  socket_t(context_t&, int) {}
};
}  // namespace zmq
#endif
