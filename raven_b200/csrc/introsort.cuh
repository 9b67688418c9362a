// raven_b200 — libstdc++'s std::sort, restated step for step.
//
// The reference truncates every overlap list with
//     std::sort(list.begin(), list.end(), by GetOverlapLength, descending)
// and keeps the first kMaxNumOverlaps (RavenLib/src/construct.cc:98-107).
// The key is not unique and std::sort is unstable, so WHICH overlaps survive
// depends on the exact sequence of moves libstdc++ makes. To keep the kept set
// bit-identical while doing the truncation on the GPU (one thread per read),
// this header restates that sequence: introsort = median-of-3 quicksort with
// an unguarded partition, depth limit 2*floor(log2 n) with a heapsort
// fallback, threshold 16, then a final (un)guarded insertion sort
// (GCC 13 <bits/stl_algo.h>, <bits/stl_heap.h>).
//
// It is compiled for the host as well (tests/test_introsort.py runs it against
// the real std::sort over random, tie-heavy and adversarial inputs).
//
// Elements are u64 = key << 32 | payload; ONLY the key takes part in
// comparisons: comp(a, b) := key(a) > key(b)   (descending, like the reference)
#pragma once

#include <cstdint>

#if defined(__CUDACC__)
#define RVN_HD __host__ __device__ __forceinline__
#else
#define RVN_HD inline
#endif

namespace rvn {
namespace stdsort {

using Elem = std::uint64_t;

RVN_HD bool Comp(Elem a, Elem b) { return (a >> 32) > (b >> 32); }

RVN_HD void Swap(Elem* a, Elem* b) {
  Elem t = *a;
  *a = *b;
  *b = t;
}

RVN_HD int Lg(std::int64_t n) {  // floor(log2 n), n > 0
  int r = 0;
  while (n > 1) {
    n >>= 1;
    ++r;
  }
  return r;
}

RVN_HD void MoveMedianToFirst(Elem* result, Elem* a, Elem* b, Elem* c) {
  if (Comp(*a, *b)) {
    if (Comp(*b, *c)) {
      Swap(result, b);
    } else if (Comp(*a, *c)) {
      Swap(result, c);
    } else {
      Swap(result, a);
    }
  } else if (Comp(*a, *c)) {
    Swap(result, a);
  } else if (Comp(*b, *c)) {
    Swap(result, c);
  } else {
    Swap(result, b);
  }
}

RVN_HD Elem* UnguardedPartition(Elem* first, Elem* last, Elem* pivot) {
  while (true) {
    while (Comp(*first, *pivot)) ++first;
    --last;
    while (Comp(*pivot, *last)) --last;
    if (!(first < last)) return first;
    Swap(first, last);
    ++first;
  }
}

RVN_HD Elem* UnguardedPartitionPivot(Elem* first, Elem* last) {
  Elem* mid = first + (last - first) / 2;
  MoveMedianToFirst(first, first + 1, mid, last - 1);
  return UnguardedPartition(first + 1, last, first);
}

// ---- heap fallback (std::__partial_sort(first, last, last)) ----
RVN_HD void PushHeap(Elem* first, std::int64_t hole, std::int64_t top, Elem value) {
  std::int64_t parent = (hole - 1) / 2;
  while (hole > top && Comp(first[parent], value)) {
    first[hole] = first[parent];
    hole = parent;
    parent = (hole - 1) / 2;
  }
  first[hole] = value;
}

RVN_HD void AdjustHeap(Elem* first, std::int64_t hole, std::int64_t len, Elem value) {
  const std::int64_t top = hole;
  std::int64_t child = hole;
  while (child < (len - 1) / 2) {
    child = 2 * (child + 1);
    if (Comp(first[child], first[child - 1])) --child;
    first[hole] = first[child];
    hole = child;
  }
  if ((len & 1) == 0 && child == (len - 2) / 2) {
    child = 2 * (child + 1);
    first[hole] = first[child - 1];
    hole = child - 1;
  }
  PushHeap(first, hole, top, value);
}

RVN_HD void HeapSortAll(Elem* first, Elem* last) {
  const std::int64_t len = last - first;
  if (len >= 2) {  // make_heap
    std::int64_t parent = (len - 2) / 2;
    while (true) {
      Elem value = first[parent];
      AdjustHeap(first, parent, len, value);
      if (parent == 0) break;
      --parent;
    }
  }
  // heap_select's scan over [middle, last) is empty when middle == last
  while (last - first > 1) {  // sort_heap
    --last;
    Elem value = *last;  // pop_heap(first, last, last)
    *last = *first;
    AdjustHeap(first, 0, last - first, value);
  }
}

// ---- insertion sorts ----
RVN_HD void UnguardedLinearInsert(Elem* last) {
  Elem val = *last;
  Elem* next = last - 1;
  while (Comp(val, *next)) {
    *last = *next;
    last = next;
    --next;
  }
  *last = val;
}

RVN_HD void InsertionSort(Elem* first, Elem* last) {
  if (first == last) return;
  for (Elem* i = first + 1; i != last; ++i) {
    if (Comp(*i, *first)) {
      Elem val = *i;
      for (Elem* p = i; p != first; --p) *p = *(p - 1);  // move_backward
      *first = val;
    } else {
      UnguardedLinearInsert(i);
    }
  }
}

// std::sort(first, last, comp)
RVN_HD void Sort(Elem* first, Elem* last) {
  if (first == last) return;
  constexpr std::int64_t kThreshold = 16;
  // __introsort_loop: recursion on the right part, iteration on the left;
  // the explicit stack holds the deferred (last, depth) of enclosing frames
  struct Frame {
    Elem* last;
    int depth;
  };
  Frame stack[2 * 64 + 2];
  int sp = 0;
  Elem* lo = first;
  Elem* hi = last;
  int depth = Lg(last - first) * 2;
  while (true) {
    if (hi - lo > kThreshold) {
      if (depth == 0) {
        HeapSortAll(lo, hi);  // this frame is done
      } else {
        --depth;
        Elem* cut = UnguardedPartitionPivot(lo, hi);
        // recurse into [cut, hi) first, then continue this frame with [lo, cut)
        stack[sp].last = cut;
        stack[sp].depth = depth;
        // remember the left part: its `first` is the current lo
        // (stored implicitly: left parts are resumed in LIFO order and their
        //  first is the `lo` that was current when pushed)
        ++sp;
        // save lo for the resumed frame
        stack[sp].last = lo;  // slot used as "first" of the deferred left part
        stack[sp].depth = 0;
        ++sp;
        lo = cut;
        continue;
      }
    }
    // current range finished: resume the most recent deferred left part
    if (sp == 0) break;
    sp -= 2;
    hi = stack[sp].last;
    depth = stack[sp].depth;
    lo = stack[sp + 1].last;
  }
  // __final_insertion_sort
  if (last - first > kThreshold) {
    InsertionSort(first, first + kThreshold);
    for (Elem* i = first + kThreshold; i != last; ++i) UnguardedLinearInsert(i);
  } else {
    InsertionSort(first, last);
  }
}

}  // namespace stdsort
}  // namespace rvn
