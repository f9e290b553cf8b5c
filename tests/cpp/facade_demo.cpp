/*
 * facade_demo.cpp — the reference's demo call sequence
 * (RSLightFields/tests/test_fine_to_coarse.cpp:56-63, test_depth_computation_2d.cpp:56-58,
 * test_depth_computation_pile.cpp:49-51) through the C++ facade of include/rslf_b200.hpp.
 * Reads a raw float32 [V][S][U][C] stack, runs the three computers and writes the
 * result maps as raw files, which tests/test_cpp_facade.py compares with the oracle.
 * usage: facade_demo in.bin V S U C D dmin dmax outprefix [colour_table.bin (256 x 3 bytes, BGR)]
 */
#include <cstdio>
#include <cstdlib>
#include <string>

#include "rslf_b200.hpp"

using namespace rslf_b200;

static void dump(const std::string& path, const Vec<Mat>& mats)
{
    FILE* f = fopen(path.c_str(), "wb");
    for (const Mat& m : mats)
        for (int r = 0; r < m.rows; ++r) fwrite(m.data + (size_t)r * m.step, m.elemSize(), m.cols, f);
    fclose(f);
}

template <typename DataType>
static int run_all(const Vec<Mat>& epis, int D, float dmin, float dmax, const std::string& out, const unsigned char* lut)
{
    Depth1DComputer<DataType> single(epis[0], dmin, dmax, D, -1, 1.0f);      /* tests/test_depth_computation.cpp:47-48 */
    single.run();
    dump(out + "_1d_depth.bin", Vec<Mat>{single.m_best_depth_u});
    dump(out + "_1d_cd.bin", Vec<Mat>{single.m_disp_confidence_u});

    Depth1DComputer_pile<DataType> pile(epis, dmin, dmax, D, -1, 1.0f);
    pile.run();
    dump(out + "_pile_depth.bin", Vec<Mat>{pile.m_best_depth_v_u});
    dump(out + "_pile_mask.bin", Vec<Mat>{pile.m_edge_confidence_mask_v_u});

    Depth2DComputer<DataType> d2(epis, dmin, dmax, D, 1.0f);
    d2.run();
    dump(out + "_2d_depth.bin", d2.get_depths_s_v_u());
    dump(out + "_2d_valid.bin", d2.get_valid_depths_mask_s_v_u());
    dump(out + "_2d_cd.bin", d2.m_disp_confidence_s_v_u);

    FineToCoarse<DataType> ftc(epis, dmin, dmax, D, 1.0f);
    ftc.run();
    Vec<Mat> map, valid;
    ftc.get_results(map, valid);
    dump(out + "_ftc_map.bin", map);
    dump(out + "_ftc_valid.bin", valid);
    if (lut) {                       /* FineToCoarse::get_coloured_depth_maps (tests/test_fine_to_coarse.cpp:71) */
        Vec<Mat> plots;
        ftc.get_coloured_depth_maps(plots, 2, true, lut);
        dump(out + "_ftc_bgr.bin", plots);
    }
    /* the free functions with their Mat signatures (core.hpp:236-375, rslf_fine_to_coarse_core.hpp:28-49) */
    {
        Depth1DParameters<DataType> prm;
        const Vec<Mat>& nepis = d2.get_epis();                       /* the normalised EPIs (dc.hpp:207) */
        dump(out + "_epis.bin", nepis);
        Mat ce, mask, med;
        compute_1D_edge_confidence_pile<DataType>(nepis, 1, ce, mask, prm);
        dump(out + "_fn_ce.bin", Vec<Mat>{ce});
        dump(out + "_fn_mask.bin", Vec<Mat>{mask});
        selective_median_filter<DataType>(d2.m_best_depth_s_v_u[2], med, nepis, 2, 5, d2.m_edge_confidence_mask_s_v_u[2], 0.1f);
        dump(out + "_fn_median.bin", Vec<Mat>{med});
        Vec<Mat> half;
        downsample_EPIs(nepis, half);
        dump(out + "_fn_down.bin", half);
        /* a two-level pyramid by hand: level 1 = the 2D computer on the downsampled EPIs */
        Depth2DComputer<DataType> d2b(half, dmin, dmax, D, 1.0f);
        d2b.run();
        d2b.set_accept_all(true);
        Vec<Vec<Mat>> disp{d2.get_depths_s_v_u(), d2b.get_depths_s_v_u()};
        Vec<Vec<Mat>> val{d2.get_valid_depths_mask_s_v_u(), d2b.get_valid_depths_mask_s_v_u()};
        Vec<Mat> fmap, fval;
        fuse_disp_maps(disp, val, fmap, fval);
        dump(out + "_fn_fuse_map.bin", fmap);
        dump(out + "_fn_fuse_valid.bin", fval);
        dump(out + "_fn_l1_depth.bin", d2b.get_depths_s_v_u());
        /* compute_2D_depth_epi with constant per-pixel bounds = the 2D computer */
        Vec<Mat> lo(d2.get_depths_s_v_u().size()), hi(lo.size()), e1, e2, e3, e4, e5, e6;
        for (size_t i = 0; i < lo.size(); ++i) {
            lo[i] = Mat(nepis.size(), nepis[0].cols, CV_32FC1); hi[i] = Mat(nepis.size(), nepis[0].cols, CV_32FC1);
            for (int r = 0; r < lo[i].rows; ++r) for (int c = 0; c < lo[i].cols; ++c) { lo[i].template at<float>(r, c) = dmin; hi[i].template at<float>(r, c) = dmax; }
        }
        compute_2D_depth_epi<DataType>(nepis, lo, hi, D, e1, e2, e3, e4, e5, e6, prm);
        dump(out + "_fn_2d_depth.bin", e5);
    }
    rslf_timing t = ftc.get_timing();
    printf("fine-to-coarse: %d levels, %d passes, %.0f pixels, %.3f ms on device\n", t.levels, t.passes, t.computed_pixels, t.ms_total);

    /* error behaviour: dim_d < 2 is rejected loudly */
    try {
        Depth2DComputer<DataType> bad(epis, dmin, dmax, 1, 1.0f);
        bad.run();
        printf("ERROR: dim_d = 1 was accepted\n");
        return 1;
    } catch (const Error& e) {
        printf("expected failure: %s\n", e.what());
    }
    return 0;
}

int main(int argc, char** argv)
{
    if (argc < 10) { fprintf(stderr, "usage: %s in.bin V S U C D dmin dmax outprefix\n", argv[0]); return 2; }
    const int V = atoi(argv[2]), S = atoi(argv[3]), U = atoi(argv[4]), C = atoi(argv[5]), D = atoi(argv[6]);
    const float dmin = (float)atof(argv[7]), dmax = (float)atof(argv[8]);
    Vec<Mat> epis(V);
    FILE* f = fopen(argv[1], "rb");
    if (!f) { perror("open"); return 2; }
    for (int v = 0; v < V; ++v) {
        epis[v] = Mat(S, U, CV_MAKETYPE(CV_32F, C));
        if (fread(epis[v].data, sizeof(float), (size_t)S * U * C, f) != (size_t)S * U * C) { fprintf(stderr, "short read\n"); return 2; }
    }
    fclose(f);
    unsigned char lut[768];
    bool have_lut = false;
    if (argc > 10) {
        FILE* g = fopen(argv[10], "rb");
        have_lut = g && fread(lut, 1, sizeof(lut), g) == sizeof(lut);
        if (g) fclose(g);
        if (!have_lut) { fprintf(stderr, "cannot read the colour table\n"); return 2; }
    }
    try {
        const unsigned char* l = have_lut ? lut : nullptr;
        return C == 3 ? run_all<Vec3f>(epis, D, dmin, dmax, argv[9], l) : run_all<float>(epis, D, dmin, dmax, argv[9], l);
    } catch (const Error& e) {
        fprintf(stderr, "rslf_b200 error %d: %s\n", e.code, e.what());
        return 1;
    }
}
