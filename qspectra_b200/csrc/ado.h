// Closed-form enumeration of HEOM auxiliary-density-operator (ADO) indices.
// Order = reference ADO_mappings (dynamics/heom.py:92-152): hierarchy levels
// 0 .. level_cutoff-1 concatenated, each level lexicographically ascending in
// the row-major flattened (site, Matsubara index) occupation vector.
#pragma once
#include <stdint.h>
#include <vector>

struct AdoTables {
    int bins, level_cutoff;
    int64_t n_ado = 0;
    int btop = 0;
    std::vector<int64_t> binom;          // Pascal triangle, btop x btop
    std::vector<int64_t> level_offset;   // first index of each level
    std::vector<uint8_t> index;          // [n_ado][bins]
    std::vector<int32_t> up, down;       // [n_ado][bins], -1 = absent

    AdoTables(int bins, int level_cutoff);
    int64_t C(int a, int b) const {
        if (b < 0 || a < 0 || b > a) return 0;
        return binom[(size_t)a * btop + b];
    }
    int64_t rank(const int *v) const;    // index of an occupation vector, -1 if outside
    void enumerate();                    // fills index/up/down
};
