// TEST INFRASTRUCTURE ONLY - headless Linux stand-in for the reference's
// template/precomp.h (which pulls in windows.h, GLFW, OpenGL, OpenCL, ImGui and
// FastNoise2 binaries that do not exist here).  It declares, in the original
// order (reference template/precomp.h:139-150,456-460), only what the render
// hot path's translation units need.  No reference code is copied: the
// reference headers are included from the patched build directory that
// oracle/ref/build_ref.py generates under oracle/_ref/ (git-ignored).
#pragma once

#include <chrono>
#include <fstream>
#include <vector>
#include <list>
#include <string>
#include <thread>
#include <memory>
#include <array>
#include <iostream>
#include <bitset>
#include <cmath>
#include <math.h>
#include <cstring>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <algorithm>
#include <assert.h>
#include <limits.h>
#include <immintrin.h>

using uchar = unsigned char;
using uint = unsigned int;
using ushort = unsigned short;
using namespace std;

#define ALIGN(x) __attribute__((aligned(x)))
#define MALLOC64(x) ((x) == 0 ? 0 : aligned_alloc(64, (((x) + 63) / 64) * 64))
#define FREE64(x) free(x)
#define CHECK_RESULT __attribute__((warn_unused_result))
#ifndef __inline
#define __inline inline
#endif

// runtime resolution (reference: compile-time SCRWIDTH/SCRHEIGHT macros,
// template/camera.h:4-5).  build_ref.py rewrites the macros to these.
extern int g_vp_scrwidth, g_vp_scrheight;

// --- GLFW key ids + input stubs (template/camera.h:115-171 only reads them) ---
#define GLFW_KEY_LEFT_SHIFT 340
#define GLFW_KEY_RIGHT_SHIFT 344
#define GLFW_KEY_LEFT_ALT 342
#define GLFW_KEY_UP 265
#define GLFW_KEY_DOWN 264
#define GLFW_KEY_LEFT 263
#define GLFW_KEY_RIGHT 262
#define GLFW_KEY_A 65
#define GLFW_KEY_D 68
#define GLFW_KEY_S 83
#define GLFW_KEY_W 87
#define GLFW_KEY_Q 81
#define GLFW_KEY_E 69
#define GLFW_KEY_R 82
#define GLFW_KEY_F 70
#define GLFW_KEY_SPACE 32
#define GLFW_KEY_ENTER 257
#define GLFW_KEY_ESCAPE 256
#define GLFW_KEY_TAB 258
#define GLFW_KEY_LEFT_CONTROL 341
#define GLFW_KEY_1 49
#define GLFW_KEY_2 50
#define GLFW_KEY_3 51
#define GLFW_KEY_4 52
inline bool IsKeyDown(const uint) { return false; }
inline bool WindowHasFocus() { return false; }
inline short GetAsyncKeyState(int) { return 0; }
inline void ExitGame() {}
inline void FatalError(const char* fmt, ...) { fprintf(stderr, "FatalError: %s\n", fmt); abort(); }

#include "imgui.h"  // no-op stub (oracle/ref/shim/imgui.h)

// minimal Surface: the path only writes screen->pixels (renderer.cpp:1878)
namespace Tmpl8
{
class Surface
{
public:
  Surface() = default;
  Surface(int w, int h) : width(w), height(h) { pixels = (uint*)MALLOC64((size_t)w * h * sizeof(uint)); ownBuffer = true; }
  ~Surface() { if (ownBuffer) FREE64(pixels); }
  uint* pixels = 0;
  int width = 0, height = 0;
  bool ownBuffer = false;
};
}
using namespace Tmpl8;

#include "tmpl8math.h"
#include <FastNoise/FastNoise.h>  // stub (binaries are Windows-only)

#include "Lighting/PointLight.h"
#include "Lighting/DirectionalLight.h"
#include "Lighting/SpotLight.h"
#include "Lighting/SkyDome.h"
#include "Lighting/SphereAreaLight.h"
#include "Materials/Material.h"
#include "BVH/BasicBVH.h"
#include "Game/PlayerCharacter.h"
#include "Game/ModifyingProp.h"

struct Timer
{
  Timer() { reseting(); }
  float elapsed() const
  {
    chrono::high_resolution_clock::time_point t2 = chrono::high_resolution_clock::now();
    chrono::duration<double> time_span = chrono::duration_cast<chrono::duration<double>>(t2 - start);
    return (float)time_span.count();
  }
  void reseting() { start = chrono::high_resolution_clock::now(); }
  chrono::high_resolution_clock::time_point start;
};

#include "common.h"

// helper for conversion of f32 colors to int: the reference's active branch
// (template/precomp.h:382-385) restated because precomp.h itself cannot be included.
inline uint RGBF32_to_RGB8(const float4* v)
{
  uint r = (uint)(255.0f * min(1.0f, v->x));
  uint g = (uint)(255.0f * min(1.0f, v->y));
  uint b = (uint)(255.0f * min(1.0f, v->z));
  return (r << 16) + (g << 8) + b;
}

class TheApp
{
public:
  virtual ~TheApp() = default;
  virtual void Init() {}
  virtual void Tick(float deltaTime) = 0;
  virtual void UI() { uiUpdated = false; }
  virtual void Shutdown() {}
  virtual void MouseUp(int) {}
  virtual void MouseDown(int) {}
  virtual void MouseMove(int, int) {}
  virtual void MouseWheel(float) {}
  virtual void KeyUp(int) {}
  virtual void KeyDown(int) {}
  Surface* screen = 0;
  bool uiUpdated;
  uint end_of_base_class = 99999;
};

#include "BVH/Shapes.h"
#include "scene.h"
#include "camera.h"
#include "renderer.h"
