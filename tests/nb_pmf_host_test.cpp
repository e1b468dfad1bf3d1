// Host build of csrc/nb_pmf.cuh (tests/test_library.py): reads "k n p" per line, prints the saddle-point pmf.
#include <stdio.h>
#include "../cellranger-atac_b200/csrc/nb_pmf.cuh"
int main() {
    double k, n, p;
    while (scanf("%lf %lf %lf", &k, &n, &p) == 3) printf("%.17g\n", nb_pmf_saddle(k, n, p, 1.0 - p));
    return 0;
}
