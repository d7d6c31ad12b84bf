// TEST INFRASTRUCTURE (oracle/ref): compiles the reference's objFunctions.cpp unmodified.
// It uses std::max/std::min (objFunctions.cpp:172-245) which collide with the min/max
// macros of scene.h:47-53, so those are undefined after the headers are in.
#include "std_first.h"
#include "ExternalLibrary/scene.h"
#include "ExternalLibrary/objects.h"
#undef min
#undef max
#include "objFunctions.cpp"
