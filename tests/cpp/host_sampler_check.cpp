// Host-side SampleSet::SampleImage (ring and rectangle forms, run-based enumeration + thinning in mct_host.cpp) against the
// naive listing of SampleSet.cpp:82-168, :180-262, :331-361 (enumerate every cell, then thin): same samples, order and RNG state.
// Built and run by tests/test_abi.py::test_host_sampler_equals_the_naive_listing (no GPU needed).

#include "mct_host.h"

#include <chrono>
#include <cstdio>
#include <random>
using namespace MultipleCameraTracking;
static int cvr(double v) { return (int)lrint(v); }
int main() {
    Matrixu m(480, 640);
    std::mt19937 g(5);
    int bad = 0, thinned = 0;
    for (int it = 0; it < 20000; it++) {
        int x = (int)(g() % 700) - 30, y = (int)(g() % 540) - 30;
        float sx = 0.8f + (g() % 100) * 0.005f, sy = 0.8f + (g() % 100) * 0.005f;
        int kind = g() % 2;
        float p0 = 1 + (g() % 400) * 0.1f, p1 = (g() % 400) * 0.1f, p2 = (g() % 300) * 0.1f, p3 = (g() % 300) * 0.1f;
        int max_n = (g() % 3 == 0) ? 100000 : 1 + g() % 200;
        uint64_t seed = g();
        // brute force
        std::vector<std::pair<int,int>> ref;
        Public::RngStream r1((int64_t)seed); Public::SetCurrentRng(&r1);
        int nrows = 480 - cvr((double)((float)80 * sy)) - 1, ncols = 640 - cvr((double)((float)40 * sx)) - 1;
        if (kind == 0) {
            if (!(p0 > p1)) continue;
            float o2 = p0 * p0, i2 = p1 * p1;
            int minrow = std::max(0, y - (int)p0), maxrow = std::min(nrows - 1, y + (int)p0), mincol = std::max(0, x - (int)p0), maxcol = std::min(ncols - 1, x + (int)p0);
            for (int r = minrow; r <= maxrow; r++) for (int c = mincol; c <= maxcol; c++) { int d = (y-r)*(y-r)+(x-c)*(x-c); if ((float)d < o2 && (float)d >= i2) ref.push_back({c, r}); }
        } else {
            float mx = p0 + p2 + 0.5f, my = p1 + p3 + 0.5f;   // max > min
            int minrow = std::max(0, y - (int)my), maxrow = std::min(nrows - 1, y + (int)my), mincol = std::max(0, x - (int)mx), maxcol = std::min(ncols - 1, x + (int)mx);
            for (int r = minrow; r <= maxrow; r++) for (int c = mincol; c <= maxcol; c++) { int dx = abs(x-c), dy = abs(y-r); if ((float)dx >= p2 || (float)dy >= p3) ref.push_back({c, r}); }
            p0 = mx; p1 = my;
        }
        float prob = ((float)(max_n + 1)) / (float)(unsigned)ref.size();
        if (prob < 1) { thinned++; size_t k = 0; for (size_t i = 0; i < ref.size(); i++) if (!(Public::randfloat() > prob)) ref[k++] = ref[i]; ref.resize(std::min(k, (size_t)max_n)); }
        Public::RngStream r2((int64_t)seed); Public::SetCurrentRng(&r2);
        Classifier::SampleSet s;
        if (kind == 0) s.SampleImage(&m, x, y, 40, 80, p0, p1, max_n, &m, nullptr, sx, sy);
        else s.SampleImage(&m, x, y, 40, 80, p0, p1, p2, p3, max_n, &m, nullptr, sx, sy);
        bool ok = s.Size() == ref.size() && r1.state == r2.state;
        for (size_t i = 0; ok && i < ref.size(); i++) ok = s[(int)i].m_col == ref[i].first && s[(int)i].m_row == ref[i].second;
        if (!ok) { if (bad < 5) printf("MISMATCH kind %d x %d y %d p %g %g %g %g max %d: %zu vs %zu\n", kind, x, y, p0, p1, p2, p3, max_n, s.Size(), ref.size()); bad++; }
    }
    printf("bad %d, thinned cases %d\n", bad, thinned);
    return bad != 0;
}
