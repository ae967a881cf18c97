// synth.cpp — seeded synthetic LiDAR sweeps and local maps.
//
// The reference ships no data (its rosbag is an external download, README.md:33-37), so tests and the
// benchmark render their own: a static world (ground plane + rotated boxes + thin vertical cylinders), a
// spinning sensor whose rays pass through the CENTRES of the range-image cells the reference would bin them
// into (imageProjection.cpp:79-88), Gaussian range noise along the ray, random drop-outs, points emitted
// column-major (column outer, ring inner) like a spinning sensor.  Host-only, no CUDA, no oracle code.
#include <cmath>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

#include "../../include/alego_b200.h"

namespace {

struct Rng {
  uint64_t s;
  explicit Rng(uint64_t seed) : s(seed * 0x9E3779B97F4A7C15ull + 0xD1B54A32D192ED03ull) {}
  uint64_t next() {
    uint64_t z = (s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
  }
  double uni() { return (next() >> 11) * (1.0 / 9007199254740992.0); }  // [0,1)
  double uni(double a, double b) { return a + (b - a) * uni(); }
  double gauss() {
    double u1 = uni(), u2 = uni();
    if (u1 < 1e-300) u1 = 1e-300;
    return std::sqrt(-2.0 * std::log(u1)) * std::cos(2.0 * M_PI * u2);
  }
};

static inline uint64_t mix(uint64_t a, uint64_t b) {
  uint64_t z = a * 0x9E3779B97F4A7C15ull ^ (b + 0x7F4A7C15ull + (a << 6) + (a >> 2));
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

struct Box { double cx, cy, hx, hy, yaw, h; };
struct Pole { double cx, cy, r, h; };

struct World {
  double ground_z;
  double extent;
  std::vector<Box> boxes;
  std::vector<Pole> poles;
};

}  // namespace

extern "C" {

void *synth_world_create(uint64_t seed, int n_boxes, int n_poles, double extent, double sensor_height) {
  World *w = new World();
  Rng rng(seed);
  w->ground_z = -sensor_height;
  w->extent = extent;
  for (int k = 0; k < n_boxes; ++k) {
    Box b;
    // keep a clear disc around the trajectory start
    do {
      b.cx = rng.uni(-extent, extent);
      b.cy = rng.uni(-extent, extent);
    } while (std::hypot(b.cx, b.cy) < 8.0);
    b.hx = rng.uni(1.0, 6.0);
    b.hy = rng.uni(1.0, 6.0);
    b.yaw = rng.uni(0, M_PI);
    b.h = rng.uni(2.0, 9.0);
    w->boxes.push_back(b);
  }
  for (int k = 0; k < n_poles; ++k) {
    Pole p;
    do {
      p.cx = rng.uni(-extent, extent);
      p.cy = rng.uni(-extent, extent);
    } while (std::hypot(p.cx, p.cy) < 4.0);
    p.r = rng.uni(0.08, 0.3);
    p.h = rng.uni(3.0, 9.0);
    w->poles.push_back(p);
  }
  return w;
}

void synth_world_destroy(void *w) { delete static_cast<World *>(w); }

// pose4 = sensor x, y, z, yaw in the world frame.  Returns the number of points written to xyzi_out
// (capacity n_scan*horizon_scan points).  jitter_cells in [0,0.5): uniform per-ray angular jitter as a
// fraction of a cell (0 = exact cell centres).
int synth_render(void *wv, const AlegoParams *P, const double *pose4, uint64_t noise_seed, double range_sigma, double dropout,
                 double max_range, double jitter_cells, float *xyzi_out) {
  const World *w = static_cast<const World *>(wv);
  const int R = P->n_scan, C = P->horizon_scan;
  const double ox = pose4[0], oy = pose4[1], oz = pose4[2], yaw = pose4[3];
  const double cyw = std::cos(yaw), syw = std::sin(yaw);
  std::vector<float> cellbuf((size_t)R * C * 4);
  std::vector<unsigned char> hitbuf((size_t)R * C, 0);
  // columns are independent (per-cell seeded noise): split them over host threads (std::thread, not OpenMP —
  // the image's default CXX has no libgomp)
  auto render_cols = [&](int c_begin, int c_end) {
  for (int c = c_begin; c < c_end; ++c) {
    for (int r = 0; r < R; ++r) {
      Rng rng(mix(mix(noise_seed, (uint64_t)r), (uint64_t)c));
      const double jr = jitter_cells > 0 ? rng.uni(-jitter_cells, jitter_cells) : 0.0;
      const double jc = jitter_cells > 0 ? rng.uni(-jitter_cells, jitter_cells) : 0.0;
      const double v = (r + jr) * P->ang_res_y - P->ang_bottom;  // centre of row r (imageProjection.cpp:80)
      const double hdeg = (c + 0.5 + jc) * P->ang_res_x;         // centre of column c (:87-88)
      const double vr = v * M_PI / 180.0, th = -hdeg * M_PI / 180.0;
      const double dsx = std::cos(vr) * std::cos(th), dsy = std::cos(vr) * std::sin(th), dsz = std::sin(vr);
      const double dx = cyw * dsx - syw * dsy, dy = syw * dsx + cyw * dsy, dz = dsz;
      double best = max_range;
      if (dz < -1e-9) {
        const double t = (w->ground_z - oz) / dz;
        if (t > 0.3 && t < best) best = t;
      }
      for (const Box &b : w->boxes) {
        const double cb = std::cos(b.yaw), sb = std::sin(b.yaw);
        const double px = cb * (ox - b.cx) + sb * (oy - b.cy), py = -sb * (ox - b.cx) + cb * (oy - b.cy), pz = oz - w->ground_z;
        const double qx = cb * dx + sb * dy, qy = -sb * dx + cb * dy, qz = dz;
        double t0 = 0.3, t1 = best;
        const double lo[3] = {-b.hx, -b.hy, 0.0}, hi[3] = {b.hx, b.hy, b.h};
        const double p3[3] = {px, py, pz}, q3[3] = {qx, qy, qz};
        bool hit = true;
        for (int a = 0; a < 3 && hit; ++a) {
          if (std::fabs(q3[a]) < 1e-12) {
            if (p3[a] < lo[a] || p3[a] > hi[a]) hit = false;
          } else {
            double ta = (lo[a] - p3[a]) / q3[a], tb = (hi[a] - p3[a]) / q3[a];
            if (ta > tb) { double s = ta; ta = tb; tb = s; }
            if (ta > t0) t0 = ta;
            if (tb < t1) t1 = tb;
            if (t0 > t1) hit = false;
          }
        }
        if (hit && t0 > 0.3 && t0 < best) best = t0;
      }
      for (const Pole &p : w->poles) {
        const double px = ox - p.cx, py = oy - p.cy;
        const double A = dx * dx + dy * dy, B = 2 * (px * dx + py * dy), Cc = px * px + py * py - p.r * p.r;
        if (A < 1e-12) continue;
        const double disc = B * B - 4 * A * Cc;
        if (disc < 0) continue;
        const double t = (-B - std::sqrt(disc)) / (2 * A);
        if (t <= 0.3 || t >= best) continue;
        const double z = oz + t * dz;
        if (z < w->ground_z || z > w->ground_z + p.h) continue;
        best = t;
      }
      const bool drop = rng.uni() < dropout;
      double noise = rng.gauss() * range_sigma;
      if (noise > 3 * range_sigma) noise = 3 * range_sigma;
      if (noise < -3 * range_sigma) noise = -3 * range_sigma;
      if (best >= max_range || drop) continue;
      const double t = best + noise;
      float *o = &cellbuf[((size_t)c * R + r) * 4];
      o[0] = (float)(dsx * t);
      o[1] = (float)(dsy * t);
      o[2] = (float)(dsz * t);
      o[3] = (float)(10.0 + 5.0 * rng.uni());
      hitbuf[(size_t)c * R + r] = 1;
    }
  }
  };
  {
    unsigned nt = std::thread::hardware_concurrency();
    if (nt == 0) nt = 1;
    if (nt > 16) nt = 16;
    std::vector<std::thread> pool;
    const int chunk = (C + (int)nt - 1) / (int)nt;
    for (unsigned k = 0; k < nt; ++k) {
      const int b = (int)k * chunk, e = b + chunk < C ? b + chunk : C;
      if (b < e) pool.emplace_back(render_cols, b, e);
    }
    for (auto &t : pool) t.join();
  }
  int n = 0;
  for (size_t k = 0; k < (size_t)R * C; ++k)
    if (hitbuf[k]) {
      std::memcpy(xyzi_out + 4 * (size_t)n, &cellbuf[k * 4], 16);
      ++n;
    }
  return n;
}

// Local map in the world frame: corner points on the vertical edges of boxes and on poles (spacing
// corner_step), surf points on the ground disc of radius `radius` and on box faces (one jittered sample per
// surf_step^2), both with isotropic noise sigma; then, if the requested counts are not reached, filler
// structure (poles / wall and ground patches) is added in the annulus [radius+10, 4*radius] — outside the
// sensor's range, so it loads the search structure without being matched.  Returns counts through n_out[2].
int synth_make_map(void *wv, uint64_t seed, int n_corner_target, int n_surf_target, double radius, double corner_step,
                   double surf_step, double sigma, float *corner_out, float *surf_out, int *n_out) {
  const World *w = static_cast<const World *>(wv);
  Rng rng(seed ^ 0xA5A5A5A5ull);
  int nc = 0, ns = 0;
  auto put = [&](float *dst, int &n, int cap, double x, double y, double z) {
    if (n >= cap) return;
    dst[4 * n + 0] = (float)(x + sigma * rng.gauss());
    dst[4 * n + 1] = (float)(y + sigma * rng.gauss());
    dst[4 * n + 2] = (float)(z + sigma * rng.gauss());
    dst[4 * n + 3] = 0.f;
    ++n;
  };
  auto edge = [&](double x, double y, double h) {
    for (double z = w->ground_z + 0.5 * corner_step * rng.uni(); z < w->ground_z + h; z += corner_step * rng.uni(0.7, 1.3))
      put(corner_out, nc, n_corner_target, x, y, z);
  };
  auto wall = [&](double x0, double y0, double x1, double y1, double h) {
    const double len = std::hypot(x1 - x0, y1 - y0);
    const int nu = (int)std::ceil(len / surf_step), nv = (int)std::ceil(h / surf_step);
    for (int a = 0; a < nu; ++a)
      for (int b = 0; b < nv; ++b) {
        const double u = (a + rng.uni()) / nu, v = (b + rng.uni()) / nv;
        put(surf_out, ns, n_surf_target, x0 + u * (x1 - x0), y0 + u * (y1 - y0), w->ground_z + v * h);
      }
  };
  auto ground_patch = [&](double cx, double cy, double half) {
    const int nu = (int)std::ceil(2 * half / surf_step);
    for (int a = 0; a < nu; ++a)
      for (int b = 0; b < nu; ++b) {
        const double x = cx - half + (a + rng.uni()) * surf_step, y = cy - half + (b + rng.uni()) * surf_step;
        put(surf_out, ns, n_surf_target, x, y, w->ground_z);
      }
  };
  for (const Box &b : w->boxes) {
    const double cb = std::cos(b.yaw), sb = std::sin(b.yaw);
    double vx[4], vy[4];
    const double sx[4] = {-1, 1, 1, -1}, sy[4] = {-1, -1, 1, 1};
    for (int k = 0; k < 4; ++k) {
      vx[k] = b.cx + cb * sx[k] * b.hx - sb * sy[k] * b.hy;
      vy[k] = b.cy + sb * sx[k] * b.hx + cb * sy[k] * b.hy;
      edge(vx[k], vy[k], b.h);
    }
    for (int k = 0; k < 4; ++k) wall(vx[k], vy[k], vx[(k + 1) & 3], vy[(k + 1) & 3], b.h);
  }
  for (const Pole &p : w->poles) edge(p.cx, p.cy, p.h);
  {  // ground disc
    const int nu = (int)std::ceil(2 * radius / surf_step);
    for (int a = 0; a < nu; ++a)
      for (int b = 0; b < nu; ++b) {
        const double x = -radius + (a + rng.uni()) * surf_step, y = -radius + (b + rng.uni()) * surf_step;
        if (x * x + y * y > radius * radius) continue;
        put(surf_out, ns, n_surf_target, x, y, w->ground_z);
      }
  }
  // filler beyond sensor range
  int guard = 0;
  while ((nc < n_corner_target || ns < n_surf_target) && guard++ < 2000000) {
    const double rr = rng.uni(radius + 10.0, 4.0 * radius), aa = rng.uni(0, 2 * M_PI);
    const double cx = rr * std::cos(aa), cy = rr * std::sin(aa);
    if (nc < n_corner_target) edge(cx, cy, rng.uni(3.0, 9.0));
    if (ns < n_surf_target) {
      if (rng.uni() < 0.5) {
        const double a2 = rng.uni(0, M_PI), len = rng.uni(4.0, 12.0);
        wall(cx, cy, cx + len * std::cos(a2), cy + len * std::sin(a2), rng.uni(2.0, 8.0));
      } else {
        ground_patch(cx, cy, rng.uni(3.0, 8.0));
      }
    }
  }
  n_out[0] = nc;
  n_out[1] = ns;
  return 0;
}

}  // extern "C"
