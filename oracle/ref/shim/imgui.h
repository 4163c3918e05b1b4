// TEST INFRASTRUCTURE ONLY - no-op ImGui stand-in so the reference's
// renderer.cpp (debug UI at renderer.cpp:2348-2971) compiles headless.
#pragma once
namespace ImGui
{
#define VP_IMGUI_NOOP(name) template <class... A> inline bool name(A&&...) { return false; }
VP_IMGUI_NOOP(Begin) VP_IMGUI_NOOP(BeginChild) VP_IMGUI_NOOP(BeginTabBar) VP_IMGUI_NOOP(BeginTabItem)
VP_IMGUI_NOOP(Button) VP_IMGUI_NOOP(Checkbox) VP_IMGUI_NOOP(CollapsingHeader) VP_IMGUI_NOOP(ColorEdit3)
VP_IMGUI_NOOP(Combo) VP_IMGUI_NOOP(DragFloat) VP_IMGUI_NOOP(End) VP_IMGUI_NOOP(EndChild)
VP_IMGUI_NOOP(EndTabBar) VP_IMGUI_NOOP(EndTabItem) VP_IMGUI_NOOP(IsItemEdited) VP_IMGUI_NOOP(SliderFloat)
VP_IMGUI_NOOP(SliderFloat3) VP_IMGUI_NOOP(SliderInt) VP_IMGUI_NOOP(Text)
#undef VP_IMGUI_NOOP
}
