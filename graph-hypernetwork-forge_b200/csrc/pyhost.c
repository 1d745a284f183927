/* pyhost.c — host-side marshalling for the drop-in `List[str]` call (HG:240, HG:264-268; SURVEY hard part 5).
 *
 * A Python list of E relation strings usually holds a few distinct OBJECTS (callers build it as
 * `[names[r] for r in rel]`).  ghf_collapse_pylist walks the list once in C, keyed on object identity: it returns,
 * for every position, the rank of its object among the distinct objects in first-occurrence order, and the position
 * of each distinct object's first occurrence.  Only those few objects are then encoded to UTF-8 and sent to the
 * device, where the content dedup (ghf_dedup_texts) runs.  Both steps keep first-occurrence order, so their
 * composition is exactly `list(dict.fromkeys(edge_texts))`.
 *
 * Long lists are walked by several threads (round 2, second session): every thread collapses its own contiguous
 * chunk into chunk-local ranks, the chunks' distinct objects are merged in chunk order (which IS first-occurrence
 * order), and a second parallel pass rewrites local ranks to global ones through a small table - no hashing in it.
 *
 * Built as its own small library (gcc -pthread, Python.h, no CUDA) and loaded with ctypes.PyDLL: the caller holds
 * the GIL for the whole call, so the list cannot change under us; the worker threads only READ the list's item
 * pointers (no reference counts, no Python API).  No Python object is created or modified here. */
#include <Python.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static inline uint64_t mix(uint64_t x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdull;
  x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull;
  x ^= x >> 33;
  return x;
}

/* open-addressing table object -> rank, growing; `order` lists the keys by rank */
typedef struct {
  PyObject** keys;
  int32_t* vals;
  uint64_t cap;
  PyObject** order;
  int64_t* first;     /* position of the first occurrence, by rank */
  int64_t count, order_cap;
} table_t;

static int table_init(table_t* t) {
  t->cap = 1024;
  t->keys = (PyObject**)calloc(t->cap, sizeof(PyObject*));
  t->vals = (int32_t*)malloc(t->cap * sizeof(int32_t));
  t->order_cap = 512;
  t->order = (PyObject**)malloc(t->order_cap * sizeof(PyObject*));
  t->first = (int64_t*)malloc(t->order_cap * sizeof(int64_t));
  t->count = 0;
  return (t->keys && t->vals && t->order && t->first) ? 0 : -3;
}
static void table_free(table_t* t) {
  free(t->keys); free(t->vals); free(t->order); free(t->first);
  memset(t, 0, sizeof(*t));
}
static int table_grow(table_t* t) {
  const uint64_t ncap = t->cap * 4;
  PyObject** nk = (PyObject**)calloc(ncap, sizeof(PyObject*));
  int32_t* nv = (int32_t*)malloc(ncap * sizeof(int32_t));
  if (!nk || !nv) { free(nk); free(nv); return -3; }
  for (uint64_t s = 0; s < t->cap; ++s)
    if (t->keys[s]) {
      uint64_t u = mix((uint64_t)(uintptr_t)t->keys[s]) & (ncap - 1);
      while (nk[u]) u = (u + 1) & (ncap - 1);
      nk[u] = t->keys[s];
      nv[u] = t->vals[s];
    }
  free(t->keys); free(t->vals);
  t->keys = nk; t->vals = nv; t->cap = ncap;
  return 0;
}
/* rank of `o`, inserted with first-occurrence position `pos` when new; -2: more than max_unique; -3: out of memory */
static inline int64_t table_rank(table_t* t, PyObject* o, int64_t pos, int64_t max_unique) {
  uint64_t s = mix((uint64_t)(uintptr_t)o) & (t->cap - 1);
  while (t->keys[s] && t->keys[s] != o) s = (s + 1) & (t->cap - 1);
  if (t->keys[s]) return t->vals[s];
  if (t->count >= max_unique) return -2;
  if (t->count == t->order_cap) {
    const int64_t nc = t->order_cap * 4;
    PyObject** no = (PyObject**)realloc(t->order, nc * sizeof(PyObject*));
    if (!no) return -3;
    t->order = no;
    int64_t* nf = (int64_t*)realloc(t->first, nc * sizeof(int64_t));
    if (!nf) return -3;
    t->first = nf;
    t->order_cap = nc;
  }
  const int64_t r = t->count++;
  t->keys[s] = o;
  t->vals[s] = (int32_t)r;
  t->order[r] = o;
  t->first[r] = pos;
  if ((uint64_t)t->count * 2 > t->cap && table_grow(t)) return -3;
  return r;
}

typedef struct {
  PyObject** items;
  int64_t a, b;            /* chunk [a, b) */
  int32_t* edge_map;
  int64_t max_unique;
  table_t local;
  const int32_t* to_global; /* pass 2: local rank -> global rank */
  int64_t status;           /* 0, -2 or -3 */
} chunk_t;

static void* chunk_pass1(void* arg) {
  chunk_t* c = (chunk_t*)arg;
  for (int64_t i = c->a; i < c->b; ++i) {
    const int64_t r = table_rank(&c->local, c->items[i], i, c->max_unique);
    if (r < 0) { c->status = r; return NULL; }
    c->edge_map[i] = (int32_t)r;
  }
  return NULL;
}
static void* chunk_pass2(void* arg) {
  chunk_t* c = (chunk_t*)arg;
  const int32_t* g = c->to_global;
  for (int64_t i = c->a; i < c->b; ++i) c->edge_map[i] = g[c->edge_map[i]];
  return NULL;
}

#define GHF_MAX_THREADS 32

/* -> number of distinct objects; -1: not a list; -2: more than max_unique distinct objects (nothing useful written);
 * -3: out of memory.  edge_map[E] int32, first[max_unique] int64.  threads <= 0: one per 1M entries, at most 16. */
int64_t ghf_collapse_pylist_mt(PyObject* list, int32_t* edge_map, int64_t* first, int64_t max_unique, int threads) {
  if (!PyList_Check(list)) return -1;
  const int64_t n = (int64_t)PyList_GET_SIZE(list);
  PyObject** items = ((PyListObject*)list)->ob_item;
  if (threads <= 0) {
    threads = (int)(n >> 20);
    if (threads > 16) threads = 16;
  }
  if (threads < 1) threads = 1;
  if (threads > GHF_MAX_THREADS) threads = GHF_MAX_THREADS;
  if (threads > n) threads = n > 0 ? (int)n : 1;

  chunk_t ch[GHF_MAX_THREADS];
  memset(ch, 0, sizeof(ch));
  int64_t rc = 0;
  for (int t = 0; t < threads; ++t) {
    ch[t].items = items;
    ch[t].a = n * t / threads;
    ch[t].b = n * (t + 1) / threads;
    ch[t].edge_map = edge_map;
    ch[t].max_unique = max_unique;
    if (table_init(&ch[t].local)) rc = -3;
  }
  pthread_t tid[GHF_MAX_THREADS];
  int started[GHF_MAX_THREADS] = {0};
  if (rc == 0) {
    for (int t = 1; t < threads; ++t) started[t] = pthread_create(&tid[t], NULL, chunk_pass1, &ch[t]) == 0;
    chunk_pass1(&ch[0]);
    for (int t = 1; t < threads; ++t) {
      if (started[t]) pthread_join(tid[t], NULL);
      else chunk_pass1(&ch[t]);                          /* could not start a thread: do its chunk here */
    }
    for (int t = 0; t < threads; ++t)
      if (ch[t].status) rc = ch[t].status;
  }
  /* merge in chunk order: chunk-local first-occurrence order, chunk after chunk, is global first-occurrence order */
  table_t global;
  memset(&global, 0, sizeof(global));
  int32_t* maps[GHF_MAX_THREADS] = {0};
  if (rc == 0 && table_init(&global)) rc = -3;
  for (int t = 0; t < threads && rc == 0; ++t) {
    const int64_t k = ch[t].local.count;
    maps[t] = (int32_t*)malloc((size_t)(k > 0 ? k : 1) * sizeof(int32_t));
    if (!maps[t]) { rc = -3; break; }
    for (int64_t j = 0; j < k; ++j) {
      const int64_t r = table_rank(&global, ch[t].local.order[j], ch[t].local.first[j], max_unique);
      if (r < 0) { rc = r; break; }
      maps[t][j] = (int32_t)r;
    }
    ch[t].to_global = maps[t];
  }
  if (rc == 0) {
    memcpy(first, global.first, (size_t)global.count * sizeof(int64_t));
    rc = global.count;
    if (threads > 1) {                                   /* one chunk: local ranks are global ranks already */
      for (int t = 1; t < threads; ++t) started[t] = pthread_create(&tid[t], NULL, chunk_pass2, &ch[t]) == 0;
      chunk_pass2(&ch[0]);
      for (int t = 1; t < threads; ++t) {
        if (started[t]) pthread_join(tid[t], NULL);
        else chunk_pass2(&ch[t]);
      }
    }
  }
  for (int t = 0; t < threads; ++t) {
    table_free(&ch[t].local);
    free(maps[t]);
  }
  if (global.keys) table_free(&global);
  return rc;
}

int64_t ghf_collapse_pylist(PyObject* list, int32_t* edge_map, int64_t* first, int64_t max_unique) {
  return ghf_collapse_pylist_mt(list, edge_map, first, max_unique, 0);
}
