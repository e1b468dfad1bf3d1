// Host build of csrc/spec_rule.cuh (tests/test_library.py): replays EM trajectories through the rule the fit kernel uses.
// stdin: "spec_iter n" then n lines "threshold crossing" (iterate 1, 2, ...), repeated per case.
// stdout per case: "first_guess first_iterate second_guess second_iterate" (-1 where there was none).
#include <stdio.h>
#include "../cellranger-atac_b200/csrc/spec_rule.cuh"
int main() {
    int spec_iter, n;
    while (scanf("%d %d", &spec_iter, &n) == 2) {
        SpecRule rule;
        long long g1 = -1, g2 = -1;
        int i1 = -1, i2 = -1;
        for (int i = 1; i <= n; i++) {
            double t, c;
            if (scanf("%lf %lf", &t, &c) != 2) return 1;
            if (!spec_rule_due(rule, i, spec_iter)) continue;
            long long g = 0;
            const int action = spec_rule_step(rule, i, spec_iter, t, c, &g);
            if (action == SPEC_FIRST_GUESS) { g1 = g; i1 = i; }
            if (action == SPEC_SECOND_GUESS) { g2 = g; i2 = i; }
        }
        printf("%lld %d %lld %d\n", g1, i1, g2, i2);
    }
    return 0;
}
