// Test stub: the reference includes this header but uses nothing from it (KMerCounter.h:17,92).
#pragma once
namespace tbb {}
