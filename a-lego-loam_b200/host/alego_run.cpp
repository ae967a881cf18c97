// alego_run — minimal offline driver of the host cores (alego_host.h): the stand-in for `roslaunch alego test.launch`
// + `rosbag play` (README.md:27-37) where ROS does not exist.  Reads a recorded batch of sweeps, runs
// ImageProjection -> LaserOdometry -> LaserMapping per sweep exactly as the nodelets are chained, writes the poses.
//
// input file (little endian): int32 preset, n_seq, n_sweeps, lm_every ; per sequence: int32 n_corner, n_surf,
// float32 corner[n_corner][4], surf[n_surf][4] ; per sweep, per sequence: int32 n, float32 xyzi[n][4]
// output file: per sweep, per sequence: float64[12] = LM params_[6], LO t_w_cur_[3], LM t_map2laser[3]
// optional 4th argument keyframe_every = k > 0: closed-loop mapping — after every k-th mapped sweep the sweep's downsampled
// clouds are stored as a keyframe at the mapped pose (saveKeyFramesAndFactor's cloud side) and the local map is re-assembled
// from the stored keyframes on the device (extractSurroundingKeyFrames, laserMapping.cpp:194-323).
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "alego_host.h"

static bool rd(FILE *f, void *dst, size_t bytes) { return bytes == 0 || std::fread(dst, 1, bytes, f) == bytes; }

int main(int argc, char **argv) {
  if (argc < 3) { std::fprintf(stderr, "usage: alego_run <sweeps.bin> <poses.bin> [device] [keyframe_every]\n"); return 2; }
  const int keyframe_every = argc > 4 ? std::atoi(argv[4]) : 0;
  FILE *fi = std::fopen(argv[1], "rb");
  if (!fi) { std::perror(argv[1]); return 2; }
  int32_t hdr[4];
  if (!rd(fi, hdr, sizeof hdr)) return 2;
  const int preset = hdr[0], n_seq = hdr[1], n_sweeps = hdr[2], lm_every = hdr[3];
  AlegoParams p;
  if (alego_default_params(&p, preset) != ALEGO_OK) { std::fprintf(stderr, "bad preset\n"); return 2; }
  alego::AlegoContext ctx;
  if (ctx.init(p, argc > 3 ? std::atoi(argv[3]) : 0, n_seq) != ALEGO_OK) {
    std::fprintf(stderr, "alego_create failed: %s\n", ctx.last_error().c_str());
    return 1;
  }
  alego::ImageProjection ip(ctx);
  alego::LaserOdometry lo(ctx);
  alego::LaserMapping lm(ctx);
  if (ip.onInit() != ALEGO_OK || lo.onInit() != ALEGO_OK || lm.onInit() != ALEGO_OK) return 1;
  for (int b = 0; b < n_seq; ++b) {
    int32_t n2[2];
    if (!rd(fi, n2, sizeof n2)) return 2;
    alego::PointCloud corner(n2[0]), surf(n2[1]);
    if (!rd(fi, corner.data(), corner.size() * 16) || !rd(fi, surf.data(), surf.size() * 16)) return 2;
    if (lm_every > 0 && lm.setLocalMap(b, corner, surf) != ALEGO_OK) { std::fprintf(stderr, "%s\n", ctx.last_error().c_str()); return 1; }
  }
  FILE *fo = std::fopen(argv[2], "wb");
  if (!fo) { std::perror(argv[2]); return 2; }
  std::vector<alego::PointCloud> clouds(n_seq);
  for (int t = 0; t < n_sweeps; ++t) {
    for (int b = 0; b < n_seq; ++b) {
      int32_t n;
      if (!rd(fi, &n, 4)) return 2;
      clouds[b].resize(n);
      if (!rd(fi, clouds[b].data(), (size_t)n * 16)) return 2;
    }
    int rc = ip.process(clouds);
    if (rc < 0) { std::fprintf(stderr, "ImageProjection: %s\n", ctx.last_error().c_str()); return 1; }
    rc = lo.process();
    if (rc < 0) { std::fprintf(stderr, "LaserOdometry: %s\n", ctx.last_error().c_str()); return 1; }
    if (lm_every > 0 && t % lm_every == 0) {  // every lm_every-th frame (reference: 2, laserMapping.cpp:112)
      rc = lm.process();
      if (rc < 0) { std::fprintf(stderr, "LaserMapping: %s\n", ctx.last_error().c_str()); return 1; }
      if (keyframe_every > 0 && (t / lm_every) % keyframe_every == 0) {
        for (int b = 0; b < n_seq; ++b) {
          double prm[6], tl[3], rl[9], to[3], ro[9];
          if (lm.pose(b, prm, tl, rl, to, ro) != ALEGO_OK) return 1;
          float pose6[6];
          for (int q = 0; q < 6; ++q) pose6[q] = (float)prm[q];  // PointTypePose fields are float (utility.h:83-97)
          if (lm.saveKeyFrame(b, pose6) != ALEGO_OK || lm.extractSurroundingKeyFrames(b) != ALEGO_OK) {
            std::fprintf(stderr, "keyframes: %s\n", ctx.last_error().c_str());
            return 1;
          }
        }
      }
    }
    for (int b = 0; b < n_seq; ++b) {
      double out[12], lop[6], rw[9], r1[9], t2[3], r2[9];
      if (lm.pose(b, out, out + 9, r1, t2, r2) != ALEGO_OK || lo.odometry(b, lop, out + 6, rw) != ALEGO_OK) return 1;
      std::fwrite(out, sizeof(double), 12, fo);
    }
  }
  std::fclose(fo);
  std::fclose(fi);
  std::printf("alego_run: %d sweeps x %d sequences done\n", n_sweeps, n_seq);
  return 0;
}
