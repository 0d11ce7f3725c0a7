/* chipmunk/chipmunk_unsafe.h -- the shape mutators of the reference's chipmunk_unsafe.h:47-60 are
 * declared in chipmunk.h in this build; this header exists so that sources including it compile. */
#ifndef CHIPMUNK_B200_UNSAFE_H
#define CHIPMUNK_B200_UNSAFE_H
#include "chipmunk.h"
#endif
