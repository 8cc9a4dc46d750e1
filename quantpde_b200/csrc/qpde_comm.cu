// qpde_comm.cu — the communicator behind the one path that has a real exchange step: the 2-D GMWB
// solve sharded by the A-axis over the GPUs of a node.  One process
// per GPU; NCCL over NVLink / NVSwitch carries the exchange on the library's own streams.
//
// libnccl is loaded at run time (dlopen of libnccl.so.2 — in a process that has imported torch that
// is the copy torch already loaded), so libqpde_b200.so has no link-time dependency on it and loads
// on a machine without NCCL; only qpde_comm_* / qpde_gmwb2d_solve_sharded need it and fail loudly
// when it is missing.  The 1-D contract batches never come here: they shard with no collective.
#include <dlfcn.h>
#include <nccl.h>

#include <cstring>
#include <string>

#include "qpde_kernels.cuh"

namespace qpde {

namespace {

struct Nccl {
	void *lib = nullptr;
	ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
	ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
	ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
	ncclResult_t (*GroupStart)() = nullptr;
	ncclResult_t (*GroupEnd)() = nullptr;
	ncclResult_t (*Broadcast)(const void *, void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
	ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
	const char *(*GetErrorString)(ncclResult_t) = nullptr;
	std::string error;
};

Nccl *nccl_api() {
	static Nccl api;
	static bool tried = false;
	if(tried) return api.lib ? &api : nullptr;
	tried = true;
	const char *names[] = { "libnccl.so.2", "libnccl.so" };
	for(const char *name : names) {
		api.lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
		if(api.lib) break;
	}
	if(!api.lib) { api.error = std::string("cannot load libnccl.so.2: ") + dlerror(); return nullptr; }
	bool ok = true;
	auto sym = [&] (const char *name) -> void * {
		void *p = dlsym(api.lib, name);
		if(!p) { ok = false; api.error = std::string("libnccl lacks ") + name; }
		return p;
	};
	api.GetUniqueId = (decltype(api.GetUniqueId))sym("ncclGetUniqueId");
	api.CommInitRank = (decltype(api.CommInitRank))sym("ncclCommInitRank");
	api.CommDestroy = (decltype(api.CommDestroy))sym("ncclCommDestroy");
	api.GroupStart = (decltype(api.GroupStart))sym("ncclGroupStart");
	api.GroupEnd = (decltype(api.GroupEnd))sym("ncclGroupEnd");
	api.Broadcast = (decltype(api.Broadcast))sym("ncclBroadcast");
	api.AllReduce = (decltype(api.AllReduce))sym("ncclAllReduce");
	api.GetErrorString = (decltype(api.GetErrorString))sym("ncclGetErrorString");
	if(!ok) { dlclose(api.lib); api.lib = nullptr; return nullptr; }
	return &api;
}

} // namespace

struct Comm {
	Nccl *api;
	ncclComm_t comm;
	int rank, world;
};

int comm_unique_id(unsigned char *id, std::string *error) {
	Nccl *api = nccl_api();
	if(!api) { *error = "NCCL unavailable"; return QPDE_ERR_UNSUPPORTED; }
	static_assert(sizeof(ncclUniqueId) == QPDE_COMM_ID_BYTES, "ncclUniqueId is 128 bytes");
	ncclUniqueId u;
	const ncclResult_t rc = api->GetUniqueId(&u);
	if(rc != ncclSuccess) { *error = std::string("ncclGetUniqueId: ") + api->GetErrorString(rc); return QPDE_ERR_CUDA; }
	memcpy(id, &u, sizeof u);
	return QPDE_OK;
}

int comm_create(const unsigned char *id, int rank, int world, Comm **out, std::string *error) {
	Nccl *api = nccl_api();
	if(!api) { *error = "NCCL unavailable"; return QPDE_ERR_UNSUPPORTED; }
	ncclUniqueId u;
	memcpy(&u, id, sizeof u);
	ncclComm_t c = nullptr;
	const ncclResult_t rc = api->CommInitRank(&c, world, u, rank);
	if(rc != ncclSuccess) { *error = std::string("ncclCommInitRank: ") + api->GetErrorString(rc); return QPDE_ERR_CUDA; }
	*out = new Comm{api, c, rank, world};
	return QPDE_OK;
}

void comm_destroy(Comm *c) {
	if(!c) return;
	c->api->CommDestroy(c->comm);
	delete c;
}

int comm_rank(const Comm *c) { return c ? c->rank : 0; }
int comm_world(const Comm *c) { return c ? c->world : 1; }

bool comm_group_start(Comm *c) { return c->api->GroupStart() == ncclSuccess; }
bool comm_group_end(Comm *c) { return c->api->GroupEnd() == ncclSuccess; }
// in place: the root's bytes at `buf` reach `buf` of every other rank
bool comm_broadcast(Comm *c, void *buf, size_t bytes, int root, cudaStream_t stream) {
	return c->api->Broadcast(buf, buf, bytes, ncclChar, root, c->comm, stream) == ncclSuccess;
}
bool comm_allreduce_max(Comm *c, int *buf, int count, cudaStream_t stream) {
	return c->api->AllReduce(buf, buf, (size_t)count, ncclInt32, ncclMax, c->comm, stream) == ncclSuccess;
}

} // namespace qpde
