// ref_wrap.cu -- TEST INFRASTRUCTURE ONLY.
// Thin __global__ entry points around the REFERENCE's own device functions,
// compiled with -I<reference>/chroma/cuda (see oracle/Makefile).  Nothing here
// restates reference logic: intersect_mesh / curand_init / curand_uniform are
// the reference's (mesh.h:45-126) and the toolkit's (curand_kernel.h) own code.
#include "mesh.h"
#include "random.h"

extern "C" {

// distance_to_mesh (mesh.h:131-155) plus the triangle index it discards.
__global__ void
ref_intersect(int nthreads, float3 *_origin, float3 *_direction, int *_last_hit,
              const Geometry *g, int *_triangle, float *_distance)
{
    __shared__ Geometry sg;
    if (threadIdx.x == 0)
        sg = *g;
    __syncthreads();

    int id = blockIdx.x*blockDim.x + threadIdx.x;
    if (id >= nthreads)
        return;
    g = &sg;

    float3 origin = _origin[id];
    float3 direction = _direction[id];
    direction /= norm(direction);

    float distance;
    int last = _last_hit ? _last_hit[id] : -1;
    int triangle_index = intersect_mesh(origin, direction, g, distance, last);
    _triangle[id] = triangle_index;
    if (triangle_index != -1)
        _distance[id] = distance;
}

// raw XORWOW words straight out of curand_init + curand (known-answer source
// for the engine's own generator): out[id*ndraw + k], state words d,v[0..4].
__global__ void
ref_rng_words(int nthreads, unsigned long long seed, unsigned long long first_stream,
              unsigned long long offset, int ndraw, unsigned int *out, unsigned int *state6)
{
    int id = blockIdx.x*blockDim.x + threadIdx.x;
    if (id >= nthreads)
        return;
    curandState s;
    curand_init(seed, first_stream + id, offset, &s);
    state6[6*id] = s.d;
    for (int j = 0; j < 5; j++)
        state6[6*id+1+j] = s.v[j];
    for (int k = 0; k < ndraw; k++)
        out[(size_t)id*ndraw + k] = curand(&s);
}

// curand_uniform floats from an existing state array (advances it)
__global__ void
ref_rng_uniform(int nthreads, curandState *s, int ndraw, float *out)
{
    int id = blockIdx.x*blockDim.x + threadIdx.x;
    if (id >= nthreads)
        return;
    curandState rng = s[id];
    for (int k = 0; k < ndraw; k++)
        out[(size_t)id*ndraw + k] = curand_uniform(&rng);
    s[id] = rng;
}

} // extern "C"
