// TEST INFRASTRUCTURE ONLY (oracle/). block_indirect_sort -> std::sort (same result).
#ifndef KMSC_ORACLE_SHIM_SORT_HPP_
#define KMSC_ORACLE_SHIM_SORT_HPP_
#include <algorithm>
namespace boost { namespace sort {
template <typename It>
void block_indirect_sort(It b, It e, int /*n_threads*/) { std::sort(b, e); }
}}
#endif
