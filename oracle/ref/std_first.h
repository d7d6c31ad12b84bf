// TEST INFRASTRUCTURE (oracle/ref): wrapper around the UNMODIFIED reference sources.
// scene.h:47-53 defines min/max macros that break libstdc++ headers included after it,
// so every translation unit that pulls the reference in includes the std headers first.
#pragma once
#include <algorithm>
#include <array>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <vector>
#include <math.h>
#include <time.h>
#include <unistd.h>
