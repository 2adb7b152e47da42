// Stand-in for the reference's tools/PerformanceCounter.h (the real one pulls in SDL). BVH.cpp:16 only
// constructs one and never reads it. TEST INFRASTRUCTURE ONLY.
#pragma once
namespace Atlas { namespace Tools { class PerformanceCounter { public: PerformanceCounter() {} }; } }
