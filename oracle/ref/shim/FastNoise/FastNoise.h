// TEST INFRASTRUCTURE ONLY - FastNoise2 ships Windows binaries only in the
// reference (lib/FastNoise2/lib/*.lib); only Scene::GenerateSomeNoise/Smoke
// (template/scene.cpp:226-356, game level, not on the render path) call it.
// This stand-in fills zeros so those functions link; no config uses them.
#pragma once
#include <memory>
#include <cstring>
namespace FastNoise
{
struct Generator
{
  void GenUniformGrid3D(float* out, int, int, int, int sx, int sy, int sz, float, int) const
  {
    std::memset(out, 0, sizeof(float) * (size_t)sx * sy * sz);
  }
};
struct Perlin : Generator {};
struct FractalFBm : Generator {};
template <class T> inline std::shared_ptr<T> New() { return std::make_shared<T>(); }
}
