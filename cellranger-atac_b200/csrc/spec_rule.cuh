// When does the threshold fit tell the host which threshold it is heading for?  (speculation)
//
// The fit (em_fast.cu) calls spec_rule_step every eighth EM iterate from spec_iter - 16 on with the integer threshold of the
// current parameters and the real-valued crossing behind it (ef_crossing).  From spec_iter on, the crossings of three check
// points give the limit the crossing converges to (Aitken: geometric convergence with ratio 0 < r < 0.95, or no movement at
// all); the limit is trusted once no integer lies between the current crossing and the limit pushed half as far again.
//   first guess : the trusted limit's threshold if there is one; nothing yet while the crossing still has more than an
//                 integer to travel (at most 64 iterates of patience); otherwise the iterate's own threshold
//   afterwards  : the first trusted limit either confirms the first guess or is the second guess; then the rule is settled
// Plain host + device code without any CUDA type, so that tests/spec_rule_host_test.cpp can replay recorded EM trajectories
// through the very same function (tests/test_library.py).
#pragma once
#include <math.h>

#ifdef __CUDACC__
#define SPEC_RULE_HD __host__ __device__
#else
#define SPEC_RULE_HD
#endif

struct SpecRule {
    double cross1 = -1.0, cross2 = -1.0;   // crossings 8 and 16 iterates ago (-1: none)
    long long shot1 = -1;                  // first guess once published
    bool settled = false;
};

enum { SPEC_NOTHING = 0, SPEC_FIRST_GUESS = 1, SPEC_SECOND_GUESS = 2 };

// is iterate i a check point?
SPEC_RULE_HD inline bool spec_rule_due(const SpecRule& st, int i, int spec_iter) {
    return !st.settled && i + 16 >= spec_iter && i <= spec_iter + 192 && ((i - spec_iter) & 7) == 0;
}

// threshold < 0: the current parameters have none (estimate_threshold returns None); crossing < 0 likewise
SPEC_RULE_HD inline int spec_rule_step(SpecRule& st, int i, int spec_iter, double threshold, double crossing, long long* guess) {
    int action = SPEC_NOTHING;
    if (i >= spec_iter) {
        bool valid = false, trusted = false;
        double limit = crossing;
        if (crossing >= 0.0 && st.cross1 >= 0.0 && st.cross2 >= 0.0) {
            const double d1 = st.cross1 - st.cross2, d2 = crossing - st.cross1;
            if (fabs(d2) < 1e-9) valid = true;
            else if (d1 != 0.0) {
                const double r = d2 / d1;
                if (r > 0.0 && r < 0.95) { limit = crossing + d2 * r / (1.0 - r); valid = true; }
            }
            if (valid) {
                const double far = limit + 0.5 * (limit - crossing) + (limit >= crossing ? 0.02 : -0.02);
                trusted = floor(fmin(crossing, far)) == floor(fmax(crossing, far));
            }
        }
        const long long raw = threshold < 0.0 ? 0 : (long long)threshold;
        const long long lim = limit < 0.0 ? 0 : (long long)floor(limit);
        if (st.shot1 < 0) {
            const bool on_its_way = valid && !trusted && fabs(limit - crossing) > 1.0 && i < spec_iter + 64;
            if (!on_its_way) {
                st.shot1 = trusted ? lim : raw;
                *guess = st.shot1;
                action = SPEC_FIRST_GUESS;
            }
        } else if (trusted) {
            if (lim != st.shot1) { *guess = lim; action = SPEC_SECOND_GUESS; }
            st.settled = true;
        }
    }
    st.cross2 = st.cross1;
    st.cross1 = crossing;
    return action;
}
