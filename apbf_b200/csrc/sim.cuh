// sim.cuh -- the whole-scene object shared by sim.cu (single GPU) and mgpu.cu (slab partition across GPUs)
#pragma once
#include "common.cuh"

struct apbf_mg_state {
	bool     enabled = false;
	int      rank = 0, world = 1;
	uint32_t lo[8][3], hi[8][3];   // brick of every rank in cell coordinates, inclusive
	uint32_t halo[3];              // halo width in cells per axis
	uint32_t n_owned = 0, n_total = 0;
};

// staging of the library-driven slab loop (apbf_sim_mg_substep, mgpu.cu): fixed-size messages, every count a device word
struct apbf_mg_loop {
	bool      ready = false;
	uint32_t  route_cap = 0;             // particles per peer and substep that may change owner
	uint32_t  halo_cap[8] = {};          // ghost records exchanged with rank r (the same number on both sides of a pair)
	uint32_t  halo_off[8] = {};          // block of rank r in send_ids / ghost_ids
	uint32_t  halo_total = 0;
	uint32_t* words = nullptr;           // device words (mgl_word)
	uint32_t* send_ids = nullptr;        // owned ids this rank provides as ghosts, block per destination
	uint32_t* ghost_ids = nullptr;       // ids of the ghosts this rank holds, block per source
	void*     send_buf[8] = {};
	void*     recv_buf[8] = {};
	size_t    buf_bytes[8] = {};
	uint64_t  exchanges = 0;
	// Peer-to-peer transport (apbf_sim_mg_p2p_export / _import): the pack kernels store straight into the receiver's staging
	// buffer over NVLink and raise a flag there; the unpack kernels wait on their own flags.  No NCCL call on the data path.
	bool      p2p = false;
	void*     arena = nullptr;            // this rank's receive side: per source rank two buffers (even / odd exchange) + the flags
	size_t    arena_bytes = 0;
	size_t    recv_off[8][2] = {};        // where messages from rank r land in `arena`
	size_t    flag_off = 0;               // uint32 flags[8][2] in `arena`: sequence number of the last complete message from r
	void*     peer_base[8] = {};          // the other ranks' arenas, opened through CUDA IPC
	void*     remote_recv[8][2] = {};     // in rank r's arena: the buffers for messages from THIS rank
	uint32_t* remote_flag[8] = {};        // in rank r's arena: the two flags for messages from THIS rank
	uint32_t* done_counter = nullptr;     // device word: CTAs of the running pack kernel that have finished their stores
	uint32_t  seq = 0;                    // exchanges so far (the same on every rank: the protocol is symmetric)
};

// One substep captured as a CUDA graph (sim.cu).  A graph is valid for exactly the host state it was captured in -- which of the
// two buffers of every list is current, the scratch allocations, the settings -- summarised in `key`.
struct apbf_sim_graph {
	uint64_t        key = 0;
	cudaGraphExec_t exec = nullptr;
	uint64_t        launches = 0;       // kernels in the graph (apbf_ctx_launch_count keeps counting)
	const uint32_t* nbr_pairs = nullptr; // the context's neighbour-structure book-keeping after the substep
	uint32_t        nbr_n_cap = 0;
	bool            nbr_valid = false, nbr_public = false;
};

struct apbf_sim {
	apbf_ctx*       ctx;
	apbf_sim_config cfg;
	apbf_fluid      fluid;
	apbf_neighbors  nb;
	float*          boxes;    // [2 * n_boxes * 4]
	float           last_dt;  // velocity_handling::mLastDeltaTime (velocity_handling.h:18)
	bool            no_fuse;
	bool            mg_fused = false; // slabs: the last search ran fused with spread_kernel_width
	bool mg_t2_tail_pending = false; // apbf_sim_mg_phase: the last apply sweep was launched in its committing form (phases 10 / 11)
	void*           nccl_comm = nullptr; // slabs: the library's own communicator (apbf_sim_mg_comm_init)
	std::vector<void*> owned;
	apbf_mg_state   mg;
	apbf_mg_loop    mgl;
	int             graphs_on = 1;       // replay captured substeps (apbf_sim_set_graphs; APBF_SIM_GRAPHS=0 turns it off)
	apbf_sim_graph  graphs[2];           // one per buffer parity
	cudaStream_t    copy_stream = nullptr;     // apbf_sim_step_host: second stream for the copies that overlap the substep
	cudaEvent_t     ev_fork = nullptr, ev_up = nullptr, ev_lists = nullptr, ev_down = nullptr;
	cudaStream_t    capture_stream = nullptr; // captures are recorded here (the caller's stream may be the legacy default stream, which cannot capture)
	uint64_t        seen_keys[4] = {};   // states met recently: a state met twice is worth a capture
	uint64_t        bad_key = 0;         // a capture of this state failed: do not try again
	uint64_t        graph_replays = 0;
	apbf_transfers  tr = {};             // cfg.transfers: the transfer list (pool.cpp:8)
	uint32_t*       sorted_index = nullptr; // cfg.transfers: the search's permutation of the hidden list, for the transfers to follow
};

extern "C" void apbf_sim_mg_comm_destroy(apbf_sim* sim); // mgpu.cu (internal: called by apbf_sim_destroy)
// data <-> reorder_out of every list (after a search or a re-partition)
void apbf_sim_swap_buffers(apbf_sim* sim);
