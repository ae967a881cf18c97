// stdsort_clone.cuh — sequential re-implementation of libstdc++'s std::sort (introsort: median-of-3
// unguarded partition, depth limit 2*floor(log2 n), heapsort fallback, final insertion sort with threshold
// 16), written from the published algorithm so that it yields THE SAME PERMUTATION as
//   std::sort(first, last, [](a, b) { return key(a) < key(b); })
// including the order it leaves between elements with equal keys.  The reference sorts each of the 6 ring
// segments with a curvature-only comparator (laserOdometry.cpp:185); equal curvatures are common (the sum is
// a float, quantised to ~1e-5), so bit-exact feature indices need the same tie order.  Elements are 64-bit
// words whose HIGH 32 bits are the key (IEEE bits of a non-negative float — monotonic as unsigned) and whose
// low 32 bits are the payload (point index); the comparator looks at the high half only.
// tests/test_stdsort_clone.py checks the clone against the real std::sort on tie-heavy inputs.
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define SSC_HD __host__ __device__ __forceinline__
#else
#define SSC_HD inline
#endif

namespace ssc {

typedef unsigned long long elem_t;
SSC_HD bool lt(elem_t a, elem_t b) { return (uint32_t)(a >> 32) < (uint32_t)(b >> 32); }

SSC_HD void swap_e(elem_t *a, elem_t *b) {
  elem_t t = *a;
  *a = *b;
  *b = t;
}

// std::__push_heap
SSC_HD void push_heap(elem_t *first, long hole, long top, elem_t value) {
  long parent = (hole - 1) / 2;
  while (hole > top && lt(first[parent], value)) {
    first[hole] = first[parent];
    hole = parent;
    parent = (hole - 1) / 2;
  }
  first[hole] = value;
}
// std::__adjust_heap
SSC_HD void adjust_heap(elem_t *first, long hole, long len, elem_t value) {
  const long top = hole;
  long child = hole;
  while (child < (len - 1) / 2) {
    child = 2 * (child + 1);
    if (lt(first[child], first[child - 1])) child--;
    first[hole] = first[child];
    hole = child;
  }
  if ((len & 1) == 0 && child == (len - 2) / 2) {
    child = 2 * (child + 1);
    first[hole] = first[child - 1];
    hole = child - 1;
  }
  push_heap(first, hole, top, value);
}
// std::__partial_sort(first, last, last) == make_heap + sort_heap
SSC_HD void heap_sort(elem_t *first, long len) {
  if (len >= 2) {
    long parent = (len - 2) / 2;
    while (true) {
      elem_t v = first[parent];
      adjust_heap(first, parent, len, v);
      if (parent == 0) break;
      parent--;
    }
  }
  long last = len;
  while (last > 1) {
    --last;
    elem_t v = first[last];
    first[last] = first[0];
    adjust_heap(first, 0, last, v);
  }
}

// std::__move_median_to_first(result, a, b, c)
SSC_HD void median_to_first(elem_t *result, elem_t *a, elem_t *b, elem_t *c) {
  if (lt(*a, *b)) {
    if (lt(*b, *c)) swap_e(result, b);
    else if (lt(*a, *c)) swap_e(result, c);
    else swap_e(result, a);
  } else if (lt(*a, *c)) swap_e(result, a);
  else if (lt(*b, *c)) swap_e(result, c);
  else swap_e(result, b);
}

// std::__unguarded_partition(first, last, pivot)
SSC_HD elem_t *unguarded_partition(elem_t *first, elem_t *last, elem_t *pivot) {
  while (true) {
    while (lt(*first, *pivot)) ++first;
    --last;
    while (lt(*pivot, *last)) --last;
    if (!(first < last)) return first;
    swap_e(first, last);
    ++first;
  }
}

SSC_HD void unguarded_linear_insert(elem_t *last) {
  elem_t val = *last;
  elem_t *next = last - 1;
  while (lt(val, *next)) {
    *last = *next;
    last = next;
    --next;
  }
  *last = val;
}
SSC_HD void insertion_sort(elem_t *first, elem_t *last) {
  if (first == last) return;
  for (elem_t *i = first + 1; i != last; ++i) {
    if (lt(*i, *first)) {
      elem_t val = *i;
      for (elem_t *p = i; p != first; --p) *p = *(p - 1);  // move_backward(first, i, i + 1)
      *first = val;
    } else {
      unguarded_linear_insert(i);
    }
  }
}

// std::sort(first, first + n).  The recursion of __introsort_loop (recurse on the right part, iterate on the
// left) is unrolled with an explicit stack of (first, last, depth) — at most one entry per level.
SSC_HD void sort(elem_t *first, long n) {
  if (n <= 1) return;
  int lg = 0;
  for (long t = n; t > 1; t >>= 1) ++lg;
  struct Frame {
    elem_t *f, *l;
    int depth;
  };
  Frame stack[72];
  int sp = 0;
  stack[sp++] = Frame{first, first + n, 2 * lg};
  while (sp > 0) {
    Frame fr = stack[--sp];
    elem_t *f = fr.f, *l = fr.l;
    int depth = fr.depth;
    while (l - f > 16) {
      if (depth == 0) {
        heap_sort(f, (long)(l - f));
        break;
      }
      --depth;
      elem_t *mid = f + (l - f) / 2;
      median_to_first(f, f + 1, mid, l - 1);
      elem_t *cut = unguarded_partition(f + 1, l, f);
      // std: __introsort_loop(cut, last, depth) first, then continue with [first, cut).  The two halves are
      // disjoint, so deferring the right half (stack) and finishing the left half first gives the same result.
      if (sp < 72) stack[sp++] = Frame{cut, l, depth};
      l = cut;
    }
  }
  // std::__final_insertion_sort
  if (n > 16) {
    insertion_sort(first, first + 16);
    for (elem_t *i = first + 16; i != first + n; ++i) unguarded_linear_insert(i);
  } else {
    insertion_sort(first, first + n);
  }
}

}  // namespace ssc
