// pdl.cuh — programmatic dependent launch (griddepcontrol) for the kernel chains of a loop body.
//
// A loop body of optimizeSet (DmsaOptimizer.h:69-144) is ~40 dependent kernels, many of them a few microseconds long, so the
// launch / block-dispatch latency between two dependent kernels (~2 us each) is a measurable share of the body.  Every
// kernel of the library starts with DMSA_PDL_ENTER(): `launch_dependents` lets the NEXT kernel of the stream be set up and
// its blocks become resident while this one still runs, `wait` then holds every thread until the PREVIOUS kernel of the
// stream has completed and its memory operations are visible.  Nothing of a kernel runs before its wait, so the
// ordering is exactly that of a plain stream; only the launch latency overlaps.  Without the launch attribute
// (dmsa_b200_set_pdl(ctx, 0), or any launch that follows a copy / memset / event wait) both instructions are no-ops.
#pragma once
#define DMSA_PDL_ENTER()                                                   \
    do {                                                                   \
        asm volatile("griddepcontrol.launch_dependents;" ::: "memory");    \
        asm volatile("griddepcontrol.wait;" ::: "memory");                 \
    } while (0)

// ptxas moves non-coherent loads (ld.global.nc: __ldg, and every load through a `const T* __restrict__` kernel parameter)
// ABOVE griddepcontrol.wait — by definition their data never changes while the kernel runs — so a kernel launched early
// would read what its predecessor has not written yet (seen in the SASS of twelve kernels, e.g. `info->R` of k_accept,
// and as wrong set counts on the GPU).  Nothing in this library is read-only across a whole chain of kernels, so no
// kernel may use the non-coherent path: `__restrict__` is dropped and `__ldg` is a plain load (both are L1-cached
// ld.global on sm_100a; profiles/ has the before / after timings).
#ifndef DMSA_KEEP_NC_LOADS
#define __restrict__
#define __ldg(p) (*(p))
#endif
