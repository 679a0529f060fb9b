// Sequential sort with the exact permutation behaviour of libstdc++'s std::sort (GCC 13,
// bits/stl_algo.h:1848-1950, bits/stl_heap.h): introsort with depth limit 2*floor(log2 n),
// median-of-3 moved to `first`, unguarded Hoare partition, threshold 16, heap-sort fallback and a final
// (partly unguarded) insertion sort.
//
// Why it exists: several reference call sites sort with comparators that have ties (e.g. chainAnchorsHits
// sorts anchors by AnchorX only, pmpfinder.cpp:2465) and std::sort is not stable, so the order of tied
// elements -- and through it the chains and cords -- is whatever this algorithm produces
// (hard part 2). Runs on one lane of the warp that owns the read.
#pragma once
#include "lnr_defs.h"

namespace lnr {

template <class T, class Less>
LNR_HD void gs_unguarded_linear_insert(T * a, int last, Less less)
{
    T val = a[last];
    int next = last - 1;
    while (less(val, a[next]))
    {
        a[last] = a[next];
        last = next;
        --next;
    }
    a[last] = val;
}

template <class T, class Less>
LNR_HD void gs_insertion_sort(T * a, int first, int last, Less less)
{
    if (first == last) return;
    for (int i = first + 1; i != last; ++i)
    {
        if (less(a[i], a[first]))
        {
            T val = a[i];
            for (int k = i; k > first; --k) a[k] = a[k - 1];
            a[first] = val;
        }
        else
            gs_unguarded_linear_insert(a, i, less);
    }
}

template <class T, class Less>
LNR_HD void gs_push_heap(T * a, int first, int hole, int top, T value, Less less)
{
    int parent = (hole - 1) / 2;
    while (hole > top && less(a[first + parent], value))
    {
        a[first + hole] = a[first + parent];
        hole = parent;
        parent = (hole - 1) / 2;
    }
    a[first + hole] = value;
}

template <class T, class Less>
LNR_HD void gs_adjust_heap(T * a, int first, int hole, int len, T value, Less less)
{
    const int top = hole;
    int child = hole;
    while (child < (len - 1) / 2)
    {
        child = 2 * (child + 1);
        if (less(a[first + child], a[first + (child - 1)])) child--;
        a[first + hole] = a[first + child];
        hole = child;
    }
    if ((len & 1) == 0 && child == (len - 2) / 2)
    {
        child = 2 * (child + 1);
        a[first + hole] = a[first + (child - 1)];
        hole = child - 1;
    }
    gs_push_heap(a, first, hole, top, value, less);
}

template <class T, class Less>
LNR_HD_COLD void gs_heap_sort(T * a, int first, int last, Less less)   // __partial_sort(first, last, last)
{
    int len = last - first;
    if (len >= 2)
    {
        int parent = (len - 2) / 2;
        while (true)
        {
            T v = a[first + parent];
            gs_adjust_heap(a, first, parent, len, v, less);
            if (parent == 0) break;
            parent--;
        }
    }
    while (last - first > 1)
    {
        --last;
        T v = a[last];
        a[last] = a[first];
        gs_adjust_heap(a, first, 0, last - first, v, less);
    }
}

template <class T, class Less>
LNR_HD void gs_move_median_to_first(T * a, int result, int ia, int ib, int ic, Less less)
{
    int pick;
    if (less(a[ia], a[ib]))
    {
        if (less(a[ib], a[ic])) pick = ib;
        else if (less(a[ia], a[ic])) pick = ic;
        else pick = ia;
    }
    else if (less(a[ia], a[ic])) pick = ia;
    else if (less(a[ib], a[ic])) pick = ic;
    else pick = ib;
    T t = a[result]; a[result] = a[pick]; a[pick] = t;
}

template <class T, class Less>
LNR_HD int gs_unguarded_partition(T * a, int first, int last, int pivot, Less less)
{
    while (true)
    {
        while (less(a[first], a[pivot])) ++first;
        --last;
        while (less(a[pivot], a[last])) --last;
        if (!(first < last)) return first;
        T t = a[first]; a[first] = a[last]; a[last] = t;
        ++first;
    }
}

// std::sort(a, a + n, less)
template <class T, class Less>
LNR_HD_COLD void gnu_sort(T * a, int n, Less less)
{
    if (n <= 0) return;
    int lg = 0;
    for (unsigned v = (unsigned)n; v > 1; v >>= 1) lg++;
    // explicit stack for the `__introsort_loop(cut, last, depth)` recursion; ranges are disjoint, so the
    // processing order does not change the result
    int st_first[72], st_last[72], st_depth[72];
    int sp = 0;
    st_first[0] = 0; st_last[0] = n; st_depth[0] = lg * 2; sp = 1;
    while (sp > 0)
    {
        --sp;
        int first = st_first[sp], last = st_last[sp], depth = st_depth[sp];
        while (last - first > 16)
        {
            if (depth == 0)
            {
                gs_heap_sort(a, first, last, less);
                break;
            }
            --depth;
            int mid = first + (last - first) / 2;
            gs_move_median_to_first(a, first, first + 1, mid, last - 1, less);
            int cut = gs_unguarded_partition(a, first + 1, last, first, less);
            st_first[sp] = cut; st_last[sp] = last; st_depth[sp] = depth; sp++;
            last = cut;
        }
    }
    if (n > 16)
    {
        gs_insertion_sort(a, 0, 16, less);
        for (int i = 16; i != n; ++i) gs_unguarded_linear_insert(a, i, less);
    }
    else
        gs_insertion_sort(a, 0, n, less);
}

}  // namespace lnr
