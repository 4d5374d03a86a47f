/* Shim for building the UNMODIFIED reference sources (test infrastructure, see oracle/build_ref.py).
 * boost::shared_ptr and its casts are the std ones.
 *
 * MCT_REF_STABLE_TIES (default build): the reference ranks weak classifiers with std::sort on the VALUE only
 * (Public.h:211-245, SortableElement::operator<), so the order of candidates with exactly equal scores is whatever the
 * standard library's std::sort happens to do -- unspecified by the C++ standard, different between the MSVC runtime of the
 * shipped VS2008 solution and libstdc++.  This header is included by Public.h after <algorithm>; with the flag set, the two
 * std::sort calls resolve to a conforming implementation that keeps equal elements in index order (std::stable_sort) and
 * ranks NaN scores last.
 * oracle/build_ref.py also builds libmct_ref_native.so without the flag (libstdc++ introsort);
 * tests/test_ref_pins_oracle.py shows the two differ only where scores tie exactly. */
#ifndef MCT_SHIM_BOOST_SHARED_PTR
#define MCT_SHIM_BOOST_SHARED_PTR
#include <memory>
#include <algorithm>
#include <numeric>
#include <list>
namespace boost {
using std::shared_ptr;
using std::static_pointer_cast;
using std::dynamic_pointer_cast;
using std::const_pointer_cast;
using std::make_shared;
}
/* The reference reports every failed assertion through abortError (Public.h:288-308), which prints and calls exit(0): a test
 * process would end with status 0 in the middle of a run.  Here exit() inside the reference sources throws instead (not a
 * std::exception, so the reference's own catch blocks do not see it); ref_harness.cpp turns it into an error return. */
struct MctRefAbort { int code; };
[[noreturn]] inline void mct_shim_exit(int code) { throw MctRefAbort{code}; }
#define exit(code) mct_shim_exit(code)
#ifdef MCT_REF_STABLE_TIES
namespace std {
/* Elements are SortableElement / SortableElementRev (Public.h:180-197): operator< compares _val only (ascending resp.
 * descending).  NaN scores rank last in both: std::sort with a comparison that is not a strict weak order (NaN) is undefined
 * behaviour in the reference, and NaN scores do occur (appearance fuser, learning rate 0, empty colour bins). */
template <class It> inline void mct_sort_ties_in_index_order(It a, It b)
{
    typedef typename std::iterator_traits<It>::value_type E;
    std::stable_sort(a, b, [](const E& x, const E& y) {
        const bool xn = x._val != x._val, yn = y._val != y._val;
        if (xn || yn) return !xn && yn;
        return x < y;
    });
}
}
#define sort mct_sort_ties_in_index_order
#endif
#endif
