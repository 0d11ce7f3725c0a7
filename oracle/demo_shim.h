/* demo_shim.h -- TEST INFRASTRUCTURE: force-included when the reference's demo sources are compiled against the drop-in's
 * include/chipmunk/chipmunk.h.  The demo headers mention one debug-draw type (cpSpace.h:270 in the reference); debug draw
 * itself is outside the hot-path scope. */
#ifndef CPB_DEMO_SHIM_H
#define CPB_DEMO_SHIM_H
typedef struct cpSpaceDebugColor { float r, g, b, a; } cpSpaceDebugColor;
#endif
