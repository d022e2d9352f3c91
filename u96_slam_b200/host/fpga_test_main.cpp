// fpga_test_main.cpp -- the reference's FPGA_TEST loop (slam/src/core/main.cpp:149-187) on the GPU drop-in:
// capture a pair, setRectImage(bank = iteration % 2), software start, receiveData, dense reprojection of the
// x4-decimated disparity (main.cpp:522-551).  Prints a checksum per frame; used by tests and INTEGRATION.md.
//   g++ -std=c++17 -O2 fpga_test_main.cpp -I../../include -L../lib -lu96stereo -Wl,-rpath,../lib -o fpga_test
#include <cstdio>
#include <cstdlib>

#include "Fpga.hpp"

static uint64_t splitmix(uint64_t &s) { uint64_t z = (s += 0x9E3779B97F4A7C15ull); z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
                                        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; return z ^ (z >> 31); }

int main(int argc, char **argv)
{
    const int frames = argc > 1 ? atoi(argv[1]) : 4;
    u96::Fpga fpga;
    if (fpga.registerOpen() != 0 || fpga.memoryOpen() != 0) { fprintf(stderr, "no CUDA device: %s\n", u96_last_cuda_error()); return 2; }
    const double sx = 640.0 / 1241, sy = 480.0 / 376;                 // StereoCameraModel.cpp:108-119
    const double P_l[12] = {718.856 * sx, 0, 607.1928 * sx, 0, 0, 718.856 * sy, 185.2157 * sy, 0, 0, 0, 1, 0};
    double P_r[12]; for (int i = 0; i < 12; i++) P_r[i] = P_l[i]; P_r[3] = -386.1448 * sx;
    uint64_t seed = 1;
    if (fpga.enableGftt(true) != 0) return 6;                           // RETURN_DATA_GFTT: fpga.c:162-172
    for (int it = 0; it < frames; it++) {
        u96::Mat8 L(480, 640), R(480, 640);
        for (int y = 0; y < 480; y++)                                   // textured pair with a 12-pixel shift
            for (int x = 0; x < 640 + 12; x++) {
                const uint8_t v = (uint8_t)(splitmix(seed) >> 56);
                if (x < 640) L.data[(size_t)y * 640 + x] = v;          // L(x) = R(x - 12)
                if (x >= 12) R.data[(size_t)y * 640 + x - 12] = v;
            }
        const int bank = it % 2;                                        // main.cpp:168
        if (fpga.setRectImage(bank, L, R) != 0) return 8;
        // the reference's setRectImage writes the bank (FPGA.cpp:236-249): the pair reads back before anything has run,
        // the disparity of this bank does not exist yet
        u96::Mat8 rl, rr; u96::Mat16 depth;
        if (fpga.receiveRectImages(bank, rl, rr) != 0 || rl.data != L.data || rr.data != R.data) return 9;
        if (it < 2 && fpga.receiveDepthMap(bank, depth) == 0) return 10;
        L.data.assign(L.data.size(), 0);                                  // the caller's images are free again
        if (fpga.startXsbl(bank) != 0) return 3;
        const int active = fpga.receiveData(rl, rr, depth);
        if (active != bank) return 4;
        const u96::Mat16 small = u96::decimateDisparity(depth, 4);     // SensorData.cpp:50-58
        std::vector<float> xyz;
        if (fpga.projectDisparityTo3D(active, P_l, P_r, 4, true, xyz) != 0) return 5;
        long long sum = 0, valid = 0, at12 = 0;
        for (short s : depth.data) { if (s >= 0) { valid++; sum += s; at12 += (s >= 11 * 16 && s <= 13 * 16); } }
        int pts = 0; for (size_t i = 0; i < xyz.size(); i += 3) pts += (xyz[i] == xyz[i]);
        // keypoint branch (main.cpp:236-243): eigen map + gftt.Max, then the threshold of generateKeypoints2 (GFTT.cpp:56-68)
        std::vector<uint16_t> eig; unsigned short maxEigen = 0;
        if (fpga.receiveEigen(active, eig, &maxEigen) != 0) return 7;
        const double thr = maxEigen * 0.01;
        int cand = 0;
        for (int y = 1; y < 479; y++) for (int x = 1; x < 639; x++) cand += ((float)eig[(size_t)y * 640 + x] >= thr);
        // generateKeypoints3D (main.cpp:250-252): every 7th candidate as a keypoint at a sub-pixel position
        std::vector<float> kp, kp3;
        for (int y = 1, c = 0; y < 479; y++) for (int x = 1; x < 639; x++)
            if ((float)eig[(size_t)y * 640 + x] >= thr && (c++ % 7) == 0) { kp.push_back(x + 0.25f); kp.push_back(y + 0.5f); }
        if (fpga.generateKeypoints3D(active, P_l, P_r, kp, kp3) != 0) return 11;
        int kp_ok = 0, kp_match = 0;
        for (size_t i = 0; i < kp.size() / 2; i++) {
            const short s = depth.at((int)kp[2 * i + 1], (int)kp[2 * i]);
            const bool good = kp3[3 * i] == kp3[3 * i];
            kp_ok += good;
            kp_match += (good == (s > 0));                                // a point exists exactly where the map holds a positive disparity
        }
        printf("kpts3d %d of %d consistent %d ; ", kp_ok, (int)(kp.size() / 2), kp_match);
        printf("frame %d bank %d valid %lld mean_disp %.3f frac_at_12px %.3f decimated %dx%d eig_max %u candidates %d points %d\n", it, active,
               valid, valid ? sum / 16.0 / valid : 0.0, valid ? (double)at12 / valid : 0.0, small.cols, small.rows, (unsigned)maxEigen, cand, pts);
    }
    return 0;
}
