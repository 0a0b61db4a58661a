/* engine_config.h — force-included (-include) ahead of every TU of the drop-in engine build
 * (engine/Makefile): the reference's search, board, GTP ... sources compiled WHERE THEY LIE,
 * plus this repository's Network implementation (engine/network_b200.cpp) in place of the
 * reference's Network.cpp + OpenCL.cpp.
 *
 * The reference picks its NN backend at compile time in config.h:21-28. config.h cannot be
 * shadowed with -I (it is included with quotes from beside each source), so its include guard
 * is pre-defined here and the switches are stated instead: USE_OPENCL ON — that is the switch
 * that turns on the reference's GPU-side search behaviour which a batched GPU evaluator needs:
 *   - asynchronous policy expansion  UCTNode.cpp:89-114 (thread_can_issue / async_scored_moves)
 *   - drain points                   UCTSearch.cpp:736-738, 891-893, 988-990 (join_outstanding_cb)
 *   - GPU-tuned search defaults      GTP.cpp:70-78, UCTSearch.cpp:42-46
 *   - cfg_gpus / cfg_rowtiles        GTP.cpp:39-42, 81-84, GTP.h:17-20
 * The OpenCL backend itself is NOT built: OpenCL.h's include guard is pre-defined too, and the
 * handful of `opencl.*` entry points the search calls (OpenCL.h:113-134) are declared below and
 * implemented on the C ABI of include/leela_b200.h in engine/network_b200.cpp.
 */
#ifndef LB2_ENGINE_CONFIG_H
#define LB2_ENGINE_CONFIG_H
#define CONFIG_INCLUDED

#define HAVE_SELECT
#define GETTIMEOFDAY
#define USE_OPTIONS
#define USE_OPENCL
#define USE_SEARCH
#define PROGRAM_NAME "Leela"
#define PROGRAM_VERSION "0.11.0"
#define MAX_CPUS 1024   // (only a clamp on --threads in the reference, GTP.cpp:67: search and netbench threads mostly sleep on their evaluation)

#include <sys/time.h>
#include <time.h>
typedef int int32;       typedef unsigned int uint32;
typedef short int16;     typedef unsigned short uint16;
typedef signed char int8; typedef unsigned char uint8;
typedef long long int int64; typedef unsigned long long int uint64;
typedef struct timeval rtime_t;

#ifdef __cplusplus
#include <atomic>
#include <string>   /* MCPolicy.h uses std::string without including <string> */

/* The lower seam (OpenCL.h:113-134) as the search sees it. Same names because the reference's
 * call sites use them; the bodies talk to the B200 evaluator. */
#define OPENCL_H_INCLUDED
class OpenCL {
public:
    void initialize();                  /* lb2_init + weight upload */
    void ensure_thread_initialized();   /* nothing per thread: the device queue is shared */
    std::string get_device_name();      /* lb2_backend_name */
    bool thread_can_issue();            /* per-thread cap on outstanding async policy requests */
    void callback_finished();
    void join_outstanding_cb();         /* lb2_drain + wait until every callback has returned */
    std::atomic<int>* get_thread_results_outstanding();
    void callback_started();            /* counted when an async request is submitted */
private:
    std::atomic<int> m_cb_outstanding{0};
};
extern OpenCL opencl;
#endif
#endif
