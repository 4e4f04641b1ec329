// Test helper: oracle/select_oracle.c's restatement of libstdc++'s std::sort against std::sort itself (order of EQUAL scores included),
// with the comparator NextGenMap uses (sortLocationScore, src/ScoreBuffer.cpp:30-32).  Built and run by tests/test_select_oracle.py.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../oracle/select_oracle.h"

struct LS {
	unsigned long long loc;
	float s;
	int orig;
};
static bool cmp(LS a, LS b) { return a.s > b.s; }

int main(int argc, char **argv) {
	const int rounds = argc > 1 ? atoi(argv[1]) : 20000;
	srand(1);
	long bad = 0;
	for (int it = 0; it < rounds; ++it) {
		const int n = 1 + rand() % (it % 50 == 0 ? 3000 : 70);
		int distinct = 1 + rand() % 6;
		if (it % 7 == 0) distinct = 1000;
		std::vector<LS> a((size_t) n);
		std::vector<sel_oracle_cand> b((size_t) n);
		const int pat = rand() % 4;
		for (int i = 0; i < n; ++i) {
			float s = (float) (rand() % distinct);
			if (pat == 1) s = (float) (i % distinct);
			if (pat == 2) s = (float) ((n - i) / 3);
			a[(size_t) i] = { (unsigned long long) i, s, i };
			b[(size_t) i].location = (unsigned long long) i;
			b[(size_t) i].score = s;
			b[(size_t) i].orig = i;
		}
		std::sort(a.begin(), a.end(), cmp);
		sel_oracle_sort(b.data(), n);
		for (int i = 0; i < n; ++i)
			if (a[(size_t) i].orig != b[(size_t) i].orig) {
				++bad;
				break;
			}
	}
	printf("%d rounds, %ld differ\n", rounds, bad);
	return bad != 0;
}
