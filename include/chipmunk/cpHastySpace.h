/* chipmunk/cpHastySpace.h -- cpHastySpace entry points (reference cpHastySpace.h:14-27).
 *
 * In the reference a hasty space farms the solver loop out to at most two pthreads
 * (src/cpHastySpace.c:446-495).  Here every space already solves on the GPU, so a hasty space is a
 * regular B200 space; the thread count caps the host threads that copy body state between the device buffers
 * and the cpBody mirrors of large spaces (host/cp_space.c host_threads). */
#ifndef CHIPMUNK_B200_HASTY_SPACE_H
#define CHIPMUNK_B200_HASTY_SPACE_H
#include "chipmunk.h"
#ifdef __cplusplus
extern "C" {
#endif
typedef struct cpHastySpace cpHastySpace;
CP_EXPORT cpSpace *cpHastySpaceNew(void);
CP_EXPORT void cpHastySpaceFree(cpSpace *space);
CP_EXPORT void cpHastySpaceSetThreads(cpSpace *space, unsigned long threads);
CP_EXPORT unsigned long cpHastySpaceGetThreads(cpSpace *space);
CP_EXPORT void cpHastySpaceStep(cpSpace *space, cpFloat dt);
#ifdef __cplusplus
}
#endif
#endif
