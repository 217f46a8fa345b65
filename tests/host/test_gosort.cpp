// Feeds key arrays to the host layer's port of Go's sort.Slice (resolv.hpp: detail::GoSortSlice) and prints the
// permutation it produces; tests/test_host_cpu.py compares it with the oracle's restatement of pdqsort_func.
// stdin: one array per line ("n k0 k1 ..."); stdout: one permutation per line.
#include <cstdio>
#include <utility>
#include <vector>

#include "../../resolv_b200/csrc/host/resolv.hpp"

int main() {
    int n;
    while (scanf("%d", &n) == 1) {
        std::vector<std::pair<double, int>> v((size_t)n);
        for (int i = 0; i < n; i++) { if (scanf("%lf", &v[(size_t)i].first) != 1) return 1; v[(size_t)i].second = i; }
        resolv::detail::GoSortSlice(v, [](const std::pair<double, int>& e) { return e.first; });
        for (int i = 0; i < n; i++) printf("%d%c", v[(size_t)i].second, i + 1 == n ? '\n' : ' ');
        if (n == 0) printf("\n");
    }
    return 0;
}
