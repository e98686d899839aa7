/*
 * oracle/ref_shim/zmq_real/zmq.h -- declarations of the part of libzmq's public C API (ZeroMQ 4.x, LGPL; the API is
 * documented at api.zeromq.org) that the reference's vendored cppzmq header and src/comms use.  TEST / BENCH
 * INFRASTRUCTURE ONLY.  The image has libzmq 4.3.5 as a shared library (bundled with pyzmq) but not its C header;
 * this file lets oracle/Makefile.full compile the reference's real ZeroMQ path and link it against that library.
 * Written from the published API: constants and struct layouts are those of libzmq 4.1 and later (zmq_msg_t is 64
 * bytes, pointer-aligned).
 */
#ifndef SLREF_ZMQ_H
#define SLREF_ZMQ_H
#include <errno.h>
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define ZMQ_VERSION_MAJOR 4
#define ZMQ_VERSION_MINOR 3
#define ZMQ_VERSION_PATCH 5
#define ZMQ_MAKE_VERSION(major, minor, patch) ((major) * 10000 + (minor) * 100 + (patch))
#define ZMQ_VERSION ZMQ_MAKE_VERSION(ZMQ_VERSION_MAJOR, ZMQ_VERSION_MINOR, ZMQ_VERSION_PATCH)

/* libzmq's own error numbers */
#define ZMQ_HAUSNUMERO 156384712
#define EFSM (ZMQ_HAUSNUMERO + 51)
#define ENOCOMPATPROTO (ZMQ_HAUSNUMERO + 52)
#define ETERM (ZMQ_HAUSNUMERO + 53)
#define EMTHREAD (ZMQ_HAUSNUMERO + 54)

int zmq_errno(void);
const char* zmq_strerror(int errnum);
void zmq_version(int* major, int* minor, int* patch);

/* context */
#define ZMQ_IO_THREADS 1
#define ZMQ_MAX_SOCKETS 2
#define ZMQ_IO_THREADS_DFLT 1
#define ZMQ_MAX_SOCKETS_DFLT 1023
void* zmq_ctx_new(void);
int zmq_ctx_term(void* context);
int zmq_ctx_shutdown(void* context);
int zmq_ctx_set(void* context, int option, int optval);
int zmq_ctx_get(void* context, int option);
int zmq_ctx_destroy(void* context);

/* messages */
typedef struct zmq_msg_t {
  unsigned char _[64] __attribute__((aligned(sizeof(void*))));
} zmq_msg_t;
typedef void(zmq_free_fn)(void* data, void* hint);
int zmq_msg_init(zmq_msg_t* msg);
int zmq_msg_init_size(zmq_msg_t* msg, size_t size);
int zmq_msg_init_data(zmq_msg_t* msg, void* data, size_t size, zmq_free_fn* ffn, void* hint);
int zmq_msg_send(zmq_msg_t* msg, void* s, int flags);
int zmq_msg_recv(zmq_msg_t* msg, void* s, int flags);
int zmq_msg_close(zmq_msg_t* msg);
int zmq_msg_move(zmq_msg_t* dest, zmq_msg_t* src);
int zmq_msg_copy(zmq_msg_t* dest, zmq_msg_t* src);
void* zmq_msg_data(zmq_msg_t* msg);
size_t zmq_msg_size(const zmq_msg_t* msg);
int zmq_msg_more(const zmq_msg_t* msg);

/* socket types */
#define ZMQ_PAIR 0
#define ZMQ_PUB 1
#define ZMQ_SUB 2
#define ZMQ_REQ 3
#define ZMQ_REP 4
#define ZMQ_DEALER 5
#define ZMQ_ROUTER 6
#define ZMQ_PULL 7
#define ZMQ_PUSH 8
#define ZMQ_XPUB 9
#define ZMQ_XSUB 10
#define ZMQ_STREAM 11

/* socket options */
#define ZMQ_AFFINITY 4
#define ZMQ_IDENTITY 5
#define ZMQ_SUBSCRIBE 6
#define ZMQ_UNSUBSCRIBE 7
#define ZMQ_RATE 8
#define ZMQ_RECOVERY_IVL 9
#define ZMQ_SNDBUF 11
#define ZMQ_RCVBUF 12
#define ZMQ_RCVMORE 13
#define ZMQ_FD 14
#define ZMQ_EVENTS 15
#define ZMQ_TYPE 16
#define ZMQ_LINGER 17
#define ZMQ_RECONNECT_IVL 18
#define ZMQ_BACKLOG 19
#define ZMQ_RECONNECT_IVL_MAX 21
#define ZMQ_MAXMSGSIZE 22
#define ZMQ_SNDHWM 23
#define ZMQ_RCVHWM 24
#define ZMQ_MULTICAST_HOPS 25
#define ZMQ_RCVTIMEO 27
#define ZMQ_SNDTIMEO 28
#define ZMQ_LAST_ENDPOINT 32
#define ZMQ_ROUTER_MANDATORY 33

/* send / recv flags */
#define ZMQ_DONTWAIT 1
#define ZMQ_SNDMORE 2

/* socket monitor events */
#define ZMQ_EVENT_CONNECTED 0x0001
#define ZMQ_EVENT_CONNECT_DELAYED 0x0002
#define ZMQ_EVENT_CONNECT_RETRIED 0x0004
#define ZMQ_EVENT_LISTENING 0x0008
#define ZMQ_EVENT_BIND_FAILED 0x0010
#define ZMQ_EVENT_ACCEPTED 0x0020
#define ZMQ_EVENT_ACCEPT_FAILED 0x0040
#define ZMQ_EVENT_CLOSED 0x0080
#define ZMQ_EVENT_CLOSE_FAILED 0x0100
#define ZMQ_EVENT_DISCONNECTED 0x0200
#define ZMQ_EVENT_MONITOR_STOPPED 0x0400
#define ZMQ_EVENT_ALL 0xFFFF
/* (the vendored cppzmq declares zmq_event_t itself for libzmq >= 4.1) */

void* zmq_socket(void* context, int type);
int zmq_close(void* s);
int zmq_setsockopt(void* s, int option, const void* optval, size_t optvallen);
int zmq_getsockopt(void* s, int option, void* optval, size_t* optvallen);
int zmq_bind(void* s, const char* addr);
int zmq_connect(void* s, const char* addr);
int zmq_unbind(void* s, const char* addr);
int zmq_disconnect(void* s, const char* addr);
int zmq_send(void* s, const void* buf, size_t len, int flags);
int zmq_recv(void* s, void* buf, size_t len, int flags);
int zmq_socket_monitor(void* s, const char* addr, int events);
int zmq_sendmsg(void* s, zmq_msg_t* msg, int flags);
int zmq_recvmsg(void* s, zmq_msg_t* msg, int flags);

/* polling */
#define ZMQ_POLLIN 1
#define ZMQ_POLLOUT 2
#define ZMQ_POLLERR 4
typedef struct zmq_pollitem_t {
  void* socket;
  int fd;
  short events;
  short revents;
} zmq_pollitem_t;
int zmq_poll(zmq_pollitem_t* items, int nitems, long timeout);

int zmq_proxy(void* frontend, void* backend, void* capture);
int zmq_proxy_steerable(void* frontend, void* backend, void* capture, void* control);

#ifdef __cplusplus
}
#endif
#endif
