/* pyhost.c — host-side marshalling for the drop-in `List[str]` call (HG:240, HG:264-268; SURVEY hard part 5).
 *
 * A Python list of E relation strings usually holds a few distinct OBJECTS (callers build it as
 * `[names[r] for r in rel]`).  ghf_collapse_pylist walks the list once in C, keyed on object identity: it returns,
 * for every position, the rank of its object among the distinct objects in first-occurrence order, and the position
 * of each distinct object's first occurrence.  Only those few objects are then encoded to UTF-8 and sent to the
 * device, where the content dedup (ghf_dedup_texts) runs.  Both steps keep first-occurrence order, so their
 * composition is exactly `list(dict.fromkeys(edge_texts))`.
 *
 * Built as its own small library (gcc, Python.h, no CUDA) and loaded with ctypes.PyDLL (the GIL is held, and the
 * list cannot change under us).  No Python object is created or modified here. */
#include <Python.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static inline uint64_t mix(uint64_t x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdull;
  x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull;
  x ^= x >> 33;
  return x;
}

/* -> number of distinct objects; -1: not a list; -2: more than max_unique distinct objects (nothing useful written);
 * -3: out of memory.  edge_map[E] int32, first[max_unique] int64. */
int64_t ghf_collapse_pylist(PyObject* list, int32_t* edge_map, int64_t* first, int64_t max_unique) {
  if (!PyList_Check(list)) return -1;
  const Py_ssize_t n = PyList_GET_SIZE(list);
  uint64_t cap = 1024;
  PyObject** keys = (PyObject**)calloc(cap, sizeof(PyObject*));
  int32_t* vals = (int32_t*)malloc(cap * sizeof(int32_t));
  if (!keys || !vals) { free(keys); free(vals); return -3; }
  int64_t count = 0;
  for (Py_ssize_t i = 0; i < n; ++i) {
    PyObject* o = PyList_GET_ITEM(list, i);
    uint64_t s = mix((uint64_t)(uintptr_t)o) & (cap - 1);
    while (keys[s] && keys[s] != o) s = (s + 1) & (cap - 1);
    if (!keys[s]) {
      if (count >= max_unique) { free(keys); free(vals); return -2; }
      keys[s] = o;
      vals[s] = (int32_t)count;
      first[count++] = (int64_t)i;
      if ((uint64_t)count * 2 > cap) {                 /* grow and rehash */
        const uint64_t ncap = cap * 4;
        PyObject** nk = (PyObject**)calloc(ncap, sizeof(PyObject*));
        int32_t* nv = (int32_t*)malloc(ncap * sizeof(int32_t));
        if (!nk || !nv) { free(nk); free(nv); free(keys); free(vals); return -3; }
        for (uint64_t t = 0; t < cap; ++t)
          if (keys[t]) {
            uint64_t u = mix((uint64_t)(uintptr_t)keys[t]) & (ncap - 1);
            while (nk[u]) u = (u + 1) & (ncap - 1);
            nk[u] = keys[t];
            nv[u] = vals[t];
          }
        free(keys); free(vals);
        keys = nk; vals = nv; cap = ncap;
        s = mix((uint64_t)(uintptr_t)o) & (cap - 1);
        while (keys[s] != o) s = (s + 1) & (cap - 1);
      }
    }
    edge_map[i] = vals[s];
  }
  free(keys); free(vals);
  return count;
}
