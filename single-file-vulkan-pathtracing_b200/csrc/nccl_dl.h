// nccl_dl.h — NCCL resolved at run time with dlopen (no link-time dependency), so libbpt loads
// in any process: inside a torch process it binds to the already-loaded bundled libnccl.so.2,
// in the C++ host app to the system one. Used only by bpt_nccl_* / bpt_allgather_image.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

int bpt_nccl_get_unique_id(uint8_t id[128], std::string* err);
int bpt_nccl_comm_init(void** comm, const uint8_t id[128], int rank, int nranks, std::string* err);
void bpt_nccl_comm_destroy(void* comm);
int bpt_nccl_allgather_f32(void* comm, const void* send, void* recv, size_t count, cudaStream_t st, std::string* err);
