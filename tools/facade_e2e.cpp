/*
 * facade_e2e.cpp — end-to-end time of the COMPILED C++ drop-in (include/rslf_b200.hpp) on ordinary, pageable Mats:
 * the reference's call sequence FineToCoarse(epis, ...) -> run() -> get_results() (tests/test_fine_to_coarse.cpp:56-63),
 * constructor upload and result download included, wall clock.  Reads a raw float32 [V][S][U][C] stack.
 * usage: facade_e2e in.bin V S U C D dmin dmax scale reps [ftc|2d]
 */
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#include "rslf_b200.hpp"

using namespace rslf_b200;

template <typename DataType>
static int run(const Vec<Mat>& epis, int D, float dmin, float dmax, float scale, int reps, bool ftc)
{
    double best = 1e30, sum = 0; double samples = 0; double checksum = 0;
    for (int r = 0; r < reps + 1; ++r) {                         /* first repetition = warm-up (allocations) */
        const auto t0 = std::chrono::steady_clock::now();
        rslf_timing t;
        if (ftc) {
            FineToCoarse<DataType> f(epis, dmin, dmax, D, scale);
            f.run();
            Vec<Mat> map, valid;
            f.get_results(map, valid);
            t = f.get_timing();
            checksum = map[map.size() / 2].template at<float>(map[0].rows / 2, map[0].cols / 2);
        } else {
            Depth2DComputer<DataType> c(epis, dmin, dmax, D, scale);
            c.run();
            t = c.get_timing();
            checksum = c.m_best_depth_s_v_u[c.m_best_depth_s_v_u.size() / 2].template at<float>(1, 1);
        }
        const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        if (r > 0) { best = ms < best ? ms : best; sum += ms; }
        samples = t.samples;
    }
    printf("{\"facade_e2e_ms_mean\": %.2f, \"facade_e2e_ms_best\": %.2f, \"samples\": %.6g, \"samples_per_s\": %.6g, \"reps\": %d, \"checksum\": %.8g, "
           "\"path\": \"compiled include/rslf_b200.hpp (%s), pageable Mats, object construction + run + get_results per repetition\"}\n",
           sum / reps, best, samples, samples / (sum / reps * 1e-3), reps, checksum,
#ifdef RSLF_B200_HAVE_OPENCV
           "cv::Mat"
#else
           "stand-in Mat"
#endif
    );
    return 0;
}

int main(int argc, char** argv)
{
    if (argc < 11) { fprintf(stderr, "usage: %s in.bin V S U C D dmin dmax scale reps [ftc|2d]\n", argv[0]); return 2; }
    const int V = atoi(argv[2]), S = atoi(argv[3]), U = atoi(argv[4]), C = atoi(argv[5]), D = atoi(argv[6]);
    const float dmin = (float)atof(argv[7]), dmax = (float)atof(argv[8]), scale = (float)atof(argv[9]);
    const int reps = atoi(argv[10]);
    const bool ftc = argc < 12 || std::string(argv[11]) == "ftc";
    Vec<Mat> epis(V);
    FILE* f = fopen(argv[1], "rb");
    if (!f) { perror("open"); return 2; }
    for (int v = 0; v < V; ++v) {
        epis[v] = Mat(S, U, CV_MAKETYPE(CV_32F, C));             /* ordinary (pageable) image storage, like a cv::Mat */
        if (fread(epis[v].data, sizeof(float), (size_t)S * U * C, f) != (size_t)S * U * C) { fprintf(stderr, "short read\n"); return 2; }
    }
    fclose(f);
    try {
        return C == 3 ? run<Vec3f>(epis, D, dmin, dmax, scale, reps, ftc) : run<float>(epis, D, dmin, dmax, scale, reps, ftc);
    } catch (const Error& e) {
        fprintf(stderr, "rslf_b200 error %d: %s\n", e.code, e.what());
        return 1;
    }
}
