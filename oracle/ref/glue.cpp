// TEST INFRASTRUCTURE (oracle/ref): the reference's own unity build, minus the GLUT viewport.
// Mirrors the include order of /root/reference/main.cpp:1-11 and defines the 11 globals of
// main.cpp:17-27.  Nothing here restates reference logic; it only makes the reference link.
//
// Ray counting: the root-level Trace()/ShadowTrace() entry points are renamed while
// RenderFunctions.cpp is compiled, and re-exported below as thin counting wrappers, so that
// calls from mtlFunctions.cpp / lightFunctions.cpp (separate TUs) are counted without
// touching the reference sources.  Recursive calls inside Trace stay un-counted (they call
// the renamed symbol), which is exactly the "ray = one root-level Trace/ShadowTrace call"
// definition of SURVEY.md section 8(d).
#include "std_first.h"
#include "ExternalLibrary/scene.h"
#include "ExternalLibrary/objects.h"
#include "ExternalLibrary/materials.h"
#include "ExternalLibrary/lights.h"
#include "ExternalLibrary/xmlload.cpp"
#include "ExternalLibrary/lodepng.cpp"
#include "ExternalLibrary/cyPhotonMap.h"

#define Trace RefTraceImpl
#define ShadowTrace RefShadowTraceImpl
#include "RenderFunctions.cpp"
#undef Trace
#undef ShadowTrace

RenderImage renderImage;
Camera camera;
Sphere theSphere;
Plane thePlane;
Node rootNode;
MaterialList materials;
LightList lights;
ObjFileList objList;
TexturedColor background;
TexturedColor environment;
TextureList textureList;

thread_local unsigned long long g_traceCalls = 0;
thread_local unsigned long long g_shadowCalls = 0;

bool Trace(const Ray &r, Node *currentNode, HitInfo &hInfo)
{
    g_traceCalls++;
    return RefTraceImpl(r, currentNode, hInfo);
}

bool ShadowTrace(const Ray &r, Node *currentNode, HitInfo &hInfo)
{
    g_shadowCalls++;
    return RefShadowTraceImpl(r, currentNode, hInfo);
}

// accessors for file-static state of RenderFunctions.cpp that the harness needs
Point3 RefCalculateImageOrigin(float d) { return CalculateImageOrigin(d); }
Point3 RefCalculateCurrentPoint(int i, int j, float ox, float oy, Point3 o) { return CalculateCurrentPoint(i, j, ox, oy, o); }
void RefRender(PixelIterator &it) { Render(it); }
void RefMonteCarlo(LightList &l, const HitInfo &h, int x, int y, int bounces, int n) { MonteCarlo(l, h, x, y, bounces, n); }

// photon path (dead code at the reference's HEAD: main.cpp:31, RenderFunctions.cpp:139-142 are commented out)
void RefGeneratePhotonMap() { GeneratePhotonMap(); }
Color RefPhotonMapping(const Ray &r, const HitInfo &h) { return PhotonMapping(r, h); }
cyPhotonMap *RefPhotonMap() { return &pMap; }
Color RefMonteCarloPhoton(const HitInfo &h, int x, int y, int n) { return MonteCarloPhoton(h, x, y, n); }
