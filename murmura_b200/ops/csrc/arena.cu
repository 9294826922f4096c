// murmura_b200 — peer-mapped symmetric arena runtime (C++ / CUDA runtime API).
//
// Replaces the reference's ZeroMQ data plane (murmura/distributed/messaging.py:59-68,
// murmura/distributed/node_process.py:130-155,227-276): instead of pickling CPU state dicts into
// sockets, every rank cudaMalloc's one symmetric region (published parameter planes, published
// sketches, control page), exports it with cudaIpcGetMemHandle and maps every peer's region with
// cudaIpcOpenMemHandle, so any kernel can dereference any node's tile over NVLink 5 / NVSwitch.
// NCCL / Gloo is used once, to exchange the 64-byte handles.
#include <torch/extension.h>
#include <pybind11/pybind11.h>
#include <c10/cuda/CUDAGuard.h>
#include <cuda_runtime.h>
#include <string>
#include <vector>

namespace py = pybind11;

#define MB_CUDA_OK(expr)                                                                         \
    do {                                                                                         \
        cudaError_t _e = (expr);                                                                 \
        TORCH_CHECK(_e == cudaSuccess, #expr, " failed: ", cudaGetErrorString(_e));              \
    } while (0)

class PeerArena {
public:
    PeerArena(int64_t device, int64_t bytes, int64_t rank, int64_t world)
        : device_((int)device), bytes_((size_t)bytes), rank_((int)rank), world_((int)world), bases_(world, nullptr) {
        TORCH_CHECK(world >= 1 && rank >= 0 && rank < world);
        c10::cuda::CUDAGuard guard(device_);
        bytes_ = (bytes_ + (2u << 20) - 1) & ~((size_t)(2u << 20) - 1);      // 2 MiB granularity (TLB page)
        MB_CUDA_OK(cudaMalloc(&bases_[rank_], bytes_));
        MB_CUDA_OK(cudaMemset(bases_[rank_], 0, bytes_));
        MB_CUDA_OK(cudaDeviceSynchronize());
    }
    ~PeerArena() { close(); }

    py::bytes ipc_handle() {
        c10::cuda::CUDAGuard guard(device_);
        cudaIpcMemHandle_t h;
        MB_CUDA_OK(cudaIpcGetMemHandle(&h, bases_[rank_]));
        return py::bytes(reinterpret_cast<const char*>(&h), sizeof(h));
    }

    // handles[r] = bytes exported by rank r (own entry ignored)
    void open_peers(const std::vector<std::string>& handles) {
        TORCH_CHECK((int)handles.size() == world_, "need one handle per rank");
        c10::cuda::CUDAGuard guard(device_);
        for (int r = 0; r < world_; ++r) {
            if (r == rank_) continue;
            TORCH_CHECK(handles[r].size() == sizeof(cudaIpcMemHandle_t), "bad IPC handle size from rank ", r);
            cudaIpcMemHandle_t h;
            memcpy(&h, handles[r].data(), sizeof(h));
            MB_CUDA_OK(cudaIpcOpenMemHandle(&bases_[r], h, cudaIpcMemLazyEnablePeerAccess));
            opened_.push_back(r);
        }
    }

    int64_t base_ptr(int64_t rank) const { return reinterpret_cast<int64_t>(bases_.at(rank)); }
    int64_t nbytes() const { return (int64_t)bytes_; }
    int64_t rank() const { return rank_; }
    int64_t world() const { return world_; }

    // Tensor aliasing [byte_offset, …) of rank `rank`'s region (local or peer-mapped). The arena must outlive it.
    torch::Tensor view(int64_t rank, int64_t byte_offset, std::vector<int64_t> sizes, py::object dtype) {
        auto st = torch::python::detail::py_object_to_dtype(dtype);
        int64_t numel = 1;
        for (auto s : sizes) numel *= s;
        TORCH_CHECK(byte_offset >= 0 && (size_t)(byte_offset + numel * (int64_t)c10::elementSize(st)) <= bytes_, "view out of range");
        char* p = static_cast<char*>(bases_.at(rank)) + byte_offset;
        return torch::from_blob(p, sizes, torch::TensorOptions().dtype(st).device(torch::kCUDA, device_));
    }

    // Device table int64[world]: base[r] + byte_offset — what kernels receive as `const T* const*`.
    torch::Tensor ptr_table(int64_t byte_offset) {
        std::vector<int64_t> host(world_);
        for (int r = 0; r < world_; ++r)
            host[r] = bases_[r] ? reinterpret_cast<int64_t>(static_cast<char*>(bases_[r]) + byte_offset) : 0;
        return torch::tensor(host, torch::TensorOptions().dtype(torch::kInt64)).to(torch::Device(torch::kCUDA, device_));
    }

    void close() {
        if (closed_) return;
        closed_ = true;
        cudaSetDevice(device_);
        for (int r : opened_) cudaIpcCloseMemHandle(bases_[r]);
        if (bases_[rank_]) cudaFree(bases_[rank_]);
    }

private:
    int device_;
    size_t bytes_;
    int rank_, world_;
    std::vector<void*> bases_;
    std::vector<int> opened_;
    bool closed_ = false;
};

void bind_arena(py::module_& m) {
    py::class_<PeerArena>(m, "PeerArena")
        .def(py::init<int64_t, int64_t, int64_t, int64_t>(), py::arg("device"), py::arg("bytes"), py::arg("rank"), py::arg("world"))
        .def("ipc_handle", &PeerArena::ipc_handle)
        .def("open_peers", &PeerArena::open_peers)
        .def("base_ptr", &PeerArena::base_ptr)
        .def("nbytes", &PeerArena::nbytes)
        .def("rank", &PeerArena::rank)
        .def("world", &PeerArena::world)
        .def("view", &PeerArena::view)
        .def("ptr_table", &PeerArena::ptr_table)
        .def("close", &PeerArena::close);
}
