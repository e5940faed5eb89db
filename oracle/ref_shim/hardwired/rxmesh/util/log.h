// stand-in for rxmesh/util/log.h when compiling apps/VertexNormal/vertex_normal_hardwired.cuh on its own
// (oracle/ref_hardwired.cu): the logging macros of the host wrapper, which this driver never calls
#pragma once
#define RXMESH_INFO(...) ((void)0)
#define RXMESH_ERROR(...) ((void)0)
