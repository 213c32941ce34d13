// Exercises the C++ host-side mirror (include/coherence_b200.hpp) the way a caller of the reference's
// modules would: build a Render.scene, render_frame it, read the canvas; use Polygon / Sprite directly.
// Prints one line per check; tests/test_cpp_mirror.py compares the numbers with the Python path.
//   host_mirror            -> runs everything (needs a GPU)
//   host_mirror --no-gpu   -> only checks that the missing device is reported as Failure
#include <cstdio>
#include <cstring>
#include "../../include/coherence_b200.hpp"
using namespace coherence;

static unsigned long long checksum(const std::vector<uint32_t>& px) {
  unsigned long long s = 0;
  for (size_t i = 0; i < px.size(); i++) s += (unsigned long long)px[i] * (unsigned long long)(1 + i % 7);
  return s;
}

int main(int argc, char** argv) {
  if (argc > 1 && !strcmp(argv[1], "--no-gpu")) {
    try {
      Context::get();
      printf("context created\n");
    } catch (const Failure& f) {
      printf("Failure: %s\n", f.what());
    }
    return 0;
  }
  try {
    const int W = 200, H = 160;
    // Polygon + Sprite
    auto tri = Polygon::edgelist_of_subpath(Polygon::path_of_pointlist({{20.3, 20.1}, {120.7, 30.2}, {60.2, 150.9}}));
    auto sm = Polygon::shapeminshape_of_unsorted_edgelist(tri, Polygon::NonZero);
    printf("triangle shape_card %lld minshape_card %lld\n", (long long)Sprite::shape_card(sm.first), (long long)Sprite::shape_card(sm.second));
    Sprite::shape maxshape = sm.first - sm.second;
    auto op = Polygon::polygon_opacity(tri, Polygon::NonZero, maxshape);
    unsigned long long osum = 0;
    for (uint8_t v : op) osum += v;
    printf("triangle maxshape_card %lld opacity_sum %llu\n", (long long)Sprite::shape_card(maxshape), osum);
    Sprite::shape b = Sprite::box(10, 10, 50, 40);
    printf("bloat_card %lld erode_card %lld union_card %lld\n", (long long)Sprite::shape_card(Sprite::bloat(2, 3, b)), (long long)Sprite::shape_card(Sprite::erode(2, 3, b)),
           (long long)Sprite::shape_card(b | Sprite::translate_shape(30, 20, b)));
    try {
      Sprite::box(0, 0, -1, 5);
      printf("negative box accepted\n");
    } catch (const Failure& f) {
      printf("negative box: Failure\n");
    }
    // Render
    using namespace Render;
    scene s;
    s.push_back(Basic_Path(Fill::plain(Colour::dissolve(Colour::colour_of_rgba(200, 30, 30, 255), 180)), {Polygon::path_of_pointlist({{20.3, 20.1}, {120.7, 30.2}, {60.2, 150.9}})}));
    s.push_back(Filter(Monochrome, {Polygon::path_of_pointlist({{60.0, 40.0}, {150.0, 45.0}, {140.0, 120.0}, {70.0, 110.0}})}));
    scene grp;
    grp.push_back(Basic_Path(Fill::plain(Colour::colour_of_rgba(10, 200, 40, 255)), {Polygon::path_of_pointlist({{40.0, 40.0}, {190.5, 60.5}, {90.0, 150.0}})}));
    grp.push_back(Basic_CPG(Fill::gradient(20., 20., 150., 120., true, false, Colour::colour_of_rgba(255, 0, 0, 255), Colour::colour_of_rgba(0, 0, 255, 255)), Subtraction,
                            {Polygon::path_of_pointlist({{10.0, 90.0}, {180.0, 80.0}, {170.0, 150.0}, {30.0, 140.0}})},
                            {Polygon::path_of_pointlist({{60.0, 100.0}, {120.0, 100.0}, {120.0, 130.0}, {60.0, 130.0}})}));
    s.push_back(Group(grp, PreTrans(0.55)));
    s.push_back(Convolved(Convolve::mkgaussian(3), Basic_Path(Fill::plain(Colour::dissolve(Colour::black, 120)), {Polygon::path_of_pointlist({{100.0, 10.0}, {190.0, 12.0}, {185.0, 70.0}, {105.0, 66.0}})})));
    scene bg;
    bg.push_back(Primitive_Rectangle(Colour::lightgrey, 0., 0., (double)W, (double)H));
    set_canvas(W, H);
    view v(s, bg);
    render_frame(v, 0, 0, W, H);
    auto px = read_rgba(0, 0, W, H);
    printf("frame checksum %llu uncovered_card %lld\n", checksum(px), (long long)Sprite::shape_card(uncovered()));
    auto rgb = read_rgb888(0, 0, W, H);
    unsigned long long rs = 0;
    for (uint8_t c : rgb) rs += c;
    printf("rgb888 sum %llu\n", rs);
    try {
      Group(scene());
      printf("empty group accepted\n");
    } catch (const Failure& f) {
      printf("empty group: Failure\n");
    }
  } catch (const std::exception& e) {
    printf("unexpected: %s\n", e.what());
    return 1;
  }
  return 0;
}
