// slam_loop_main.cpp -- BASELINE config C5: the dense part of the reference's SLAM loop on the GPU drop-in.
//   stereo pair -> Fpga shim (xsbl -> bm on the GPU) -> x4-decimated disparity (SensorData.cpp:50-58)
//   -> projectDisparityTo3D + localTransform on the GPU (Stereo.cpp:157-182, StereoCameraModel.cpp:9-14)
//   -> pose transform, range gate, OcTree::updateNode, writeBinary  == buildOccupancyGridMap (main.cpp:495-561)
// The odometry/mapper of the reference (rtabmap-derived, needs OpenCV C++) is out of scope: poses are the known poses
// of the synthetic sequence (given on the command line as a per-frame translation step).  OctoMap is the reference's
// own vendored copy, compiled from where it lies by __graft_entry__.build() (never copied into this repository).
//
//   slam_loop <sequence.bin> <out.bt> [step_x step_y step_z]
//   sequence.bin: int32 W,H,N then N x (L[H*W], R[H*W]) u8
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include <octomap/OcTree.h>

#include "Fpga.hpp"

int main(int argc, char **argv)
{
    if (argc < 3) { fprintf(stderr, "usage: slam_loop <sequence.bin> <out.bt> [sx sy sz]\n"); return 1; }
    FILE *fp = fopen(argv[1], "rb");
    if (!fp) { perror("sequence"); return 1; }
    int32_t hdr[3];
    if (fread(hdr, sizeof(int32_t), 3, fp) != 3 || hdr[0] != 640 || hdr[1] != 480) { fprintf(stderr, "bad header\n"); return 1; }
    const int W = hdr[0], H = hdr[1], N = hdr[2];
    const float step[3] = {argc > 3 ? (float)atof(argv[3]) : 0.0f, argc > 4 ? (float)atof(argv[4]) : -0.05f, argc > 5 ? (float)atof(argv[5]) : 0.0f};

    u96::Fpga fpga;
    if (fpga.registerOpen() != 0 || fpga.memoryOpen() != 0) { fprintf(stderr, "no CUDA device: %s\n", u96_last_cuda_error()); return 2; }
    // KITTI-style projection matrices scaled to 640x480 (StereoCameraModel.cpp:108-119)
    const double sx = 640.0 / 1241, sy = 480.0 / 376;
    const double P_l[12] = {718.856 * sx, 0, 607.1928 * sx, 0, 0, 718.856 * sy, 185.2157 * sy, 0, 0, 0, 1, 0};
    double P_r[12]; for (int i = 0; i < 12; i++) P_r[i] = P_l[i]; P_r[3] = -386.1448 * sx;

    octomap::OcTree tree(0.1);                                    // main.cpp:499
    const float rangeMax_ = 5.0f, rangeMaxSqrd = rangeMax_ * rangeMax_;
    const int scale = 4;                                          // SensorData.cpp:50
    u96::Mat8 L(H, W), R(H, W), rl, rr;
    u96::Mat16 depth;
    std::vector<float> xyz;
    double checksum = 0; long long inserted = 0, finite = 0;
    for (int it = 0; it < N; it++) {
        if (fread(L.data.data(), 1, (size_t)W * H, fp) != (size_t)W * H || fread(R.data.data(), 1, (size_t)W * H, fp) != (size_t)W * H) return 3;
        const int bank = it % 2;                                  // main.cpp:168
        fpga.setRectImage(bank, L, R);
        if (fpga.startXsbl(bank) != 0) return 4;
        if (fpga.receiveData(rl, rr, depth) != bank) return 5;
        if (fpga.projectDisparityTo3D(bank, P_l, P_r, scale, true, xyz) != 0) return 6;   // incl. localTransform (main.cpp:538)
        // known pose of frame `it`: pure translation (main.cpp:506-519 builds it from the optimised graph)
        const float o14 = step[0] * it, o24 = step[1] * it, o34 = step[2] * it;
        const octomap::point3d sensorOrigin(o14, o24, o34);
        for (size_t i = 0; i + 2 < xyz.size(); i += 3) {
            if (!(std::isfinite(xyz[i]) && std::isfinite(xyz[i + 1]) && std::isfinite(xyz[i + 2]))) continue;
            finite++;
            // transformPoint(pt3d, optimized_pose) with r = identity (Stereo.cpp:185-198)
            const float px = 1.0f * xyz[i] + 0.0f * xyz[i + 1] + 0.0f * xyz[i + 2] + o14;
            const float py = 0.0f * xyz[i] + 1.0f * xyz[i + 1] + 0.0f * xyz[i + 2] + o24;
            const float pz = 0.0f * xyz[i] + 0.0f * xyz[i + 1] + 1.0f * xyz[i + 2] + o34;
            const octomap::point3d pt(px, py, pz);
            const octomap::point3d v(pt.x() - sensorOrigin.x(), pt.y() - sensorOrigin.y(), pt.z() - sensorOrigin.z());
            if (v.norm() <= rangeMaxSqrd) {                       // sic: norm against the squared range (main.cpp:544)
                octomap::OcTreeKey key;
                if (tree.coordToKeyChecked(pt, key)) { tree.updateNode(key, true); inserted++; checksum += px + 2.0 * py + 3.0 * pz; }
            }
        }
    }
    fclose(fp);
    tree.writeBinary(argv[2]);                                    // main.cpp:560
    printf("frames %d finite_points %lld inserted %lld leaf_nodes %zu checksum %.6f\n", N, finite, inserted, tree.getNumLeafNodes(), checksum);
    return 0;
}
