/* demo_driver.c -- TEST INFRASTRUCTURE.  Runs the reference's OWN demo sources (demo/Bench.c, PyramidStack.c, Chains.c,
 * Planet.c, Springies.c -- compiled unmodified from where they lie under /root/reference, never copied here) against
 * whichever Chipmunk2D library it is linked with, and prints the state of every body after N steps:
 *
 *   oracle/_ref/demo_ref   : linked with the unmodified reference (oracle/_ref/libchipmunk_ref.so)
 *   oracle/_ref/demo_b200  : compiled against include/chipmunk/chipmunk.h and linked with the drop-in
 *                            (chipmunk2d_b200/lib/libchipmunk_b200.so): cpSpaceStep runs on the GPU
 *
 * The demo framework (GLFW window, debug draw) is replaced by the stubs below; the drop-in build gets the one
 * debug-draw type the demo headers mention through demo_shim.h.  tests/test_gpu_demos.py compares the two outputs. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "chipmunk/chipmunk.h"
#include "ChipmunkDemo.h"

int ChipmunkDemoTicks = 0;
double ChipmunkDemoTime = 0.0;
cpVect ChipmunkDemoKeyboard = {0, 0};
cpVect ChipmunkDemoMouse = {0, 0};
cpBool ChipmunkDemoRightClick = cpFalse;
cpBool ChipmunkDemoRightDown = cpFalse;
char const *ChipmunkDemoMessageString = NULL;
#define GRABBABLE_MASK_BIT (1u<<31)
cpShapeFilter GRAB_FILTER = {CP_NO_GROUP, GRABBABLE_MASK_BIT, GRABBABLE_MASK_BIT};
cpShapeFilter NOT_GRABBABLE_FILTER = {CP_NO_GROUP, ~GRABBABLE_MASK_BIT, ~GRABBABLE_MASK_BIT};
void ChipmunkDemoPrintString(char const *fmt, ...){ (void)fmt; }
void ChipmunkDemoDefaultDrawImpl(cpSpace *space){ (void)space; }
void ChipmunkDemoFreeSpaceChildren(cpSpace *space){ (void)space; }
cpTransform ChipmunkDebugDrawVPMatrix = {1, 0, 0, 1, 0, 0};
float ChipmunkDebugDrawPointLineScale = 1.0f;
void ChipmunkDebugDrawCircle(cpVect pos, cpFloat angle, cpFloat radius, cpSpaceDebugColor o, cpSpaceDebugColor f){ }
void ChipmunkDebugDrawSegment(cpVect a, cpVect b, cpSpaceDebugColor color){ }
void ChipmunkDebugDrawFatSegment(cpVect a, cpVect b, cpFloat radius, cpSpaceDebugColor o, cpSpaceDebugColor f){ }
void ChipmunkDebugDrawPolygon(int count, const cpVect *verts, cpFloat radius, cpSpaceDebugColor o, cpSpaceDebugColor f){ }
void ChipmunkDebugDrawDot(cpFloat size, cpVect pos, cpSpaceDebugColor fillColor){ }
void ChipmunkDebugDrawBB(cpBB bb, cpSpaceDebugColor outlineColor){ }

extern ChipmunkDemo bench_list[];
extern int bench_count;
extern ChipmunkDemo PyramidStack, Chains, Planet, Springies;

static ChipmunkDemo *find_demo(const char *name)
{
	if(strcmp(name, "PyramidStack") == 0) return &PyramidStack;
	if(strcmp(name, "Chains") == 0) return &Chains;
	if(strcmp(name, "Planet") == 0) return &Planet;
	if(strcmp(name, "Springies") == 0) return &Springies;
	for(int i = 0; i < bench_count; i++){
		const char *n = strstr(bench_list[i].name, "- ");
		n = (n ? n + 2 : bench_list[i].name);
		if(strcmp(n, name) == 0) return &bench_list[i];
	}
	return NULL;
}

static int g_index = 0;
static void print_body(cpBody *b, void *data)
{
	cpVect p = cpBodyGetPosition(b), v = cpBodyGetVelocity(b);
	printf("%d %.17g %.17g %.17g %.17g %.17g %.17g\n", g_index++, p.x, p.y, v.x, v.y, cpBodyGetAngle(b), cpBodyGetAngularVelocity(b));
}

int main(int argc, char **argv)
{
	if(argc < 3){ fprintf(stderr, "usage: %s <demo name> <steps>\n", argv[0]); return 2; }
	ChipmunkDemo *demo = find_demo(argv[1]);
	if(!demo){ fprintf(stderr, "unknown demo %s\n", argv[1]); return 2; }
	int steps = atoi(argv[2]);
	srand(45073);                                  /* RunDemo, demo/ChipmunkDemo.c:378 */
	cpSpace *space = demo->initFunc();
	for(int i = 0; i < steps; i++){
		demo->updateFunc(space, demo->timestep);   /* the demo's own update: cpSpaceStep (+ whatever the demo does per frame) */
		ChipmunkDemoTicks++; ChipmunkDemoTime += demo->timestep;
	}
	g_index = 0;
	cpSpaceEachBody(space, print_body, NULL);
	return 0;
}
