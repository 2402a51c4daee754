/* kmerguts_shard.h -- optional hash-sharded signature table (BASELINE.json configs[4]).
 *
 * The reference has no counterpart: KmerGutsJava streams ONE table file on one thread (lookup, KGJ:944-1034).  This mode
 * exists for tables that should not be replicated on every GPU.  Every rank holds the signatures whose key hashes to it
 * (kg_shard_owner), encodes its OWN sequences exactly as the replicated path does (addKmers, KGJ:900-922), and one
 * exchange step each way stands where the reference's table stream stands:
 *
 *     route   : encode every valid 8-mer, bin it by owner                           (k_route)
 *     keys    : all-to-all of the binned keys, 5 bytes each                         (NCCL send/recv, or peer copies)
 *     answer  : each owner probes the keys it received against its shard            (k_answer; same bucket layout,
 *                                                                                    prefilter and probe loop as k_probe)
 *     replies : all-to-all of {index of the query, 16-byte payload} for the HITS only
 *     merge   : replies land at their residue positions; from there on the pipeline is the replicated one
 *               (run FSM, CALL / OTU records: gatherHits / processSetOfHits, KGJ:385-514)
 *
 * Results are bit-identical to kg_batch_run against the unsharded table: a lookup returns the payload stored under the
 * key wherever it lives, and everything after the lookup is the same code.
 *
 * Transports.  "direct" (default): every rank's exchange buffer is mapped into its peers (CUDA IPC between processes, peer
 * access inside one process); k_route stores every k-mer straight into its OWNER's buffer and k_answer every hit straight into
 * the ASKER's, so the NVLink transfers happen inside the kernels that produce the data; counts and arrival flags travel
 * in-band and NCCL only bootstraps (capacity agreement, IPC handles).  KG_SHARD_TRANSPORT=nccl (one process per GPU) or =copy
 * (local groups) selects the staged transport: bins that are moved afterwards by grouped ncclSend/ncclRecv or peer copies.
 * Every rank of a communicator must use the same setting.
 *
 * One process per GPU: kg_comm_init is collective (rank 0 makes an id with kg_comm_unique_id, the host application
 * hands it to every rank over whatever channel it has -- MPI, a file, a socket -- like ncclUniqueId).
 * Environment: KG_SHARD_CHUNKS (NCCL transport only; 1-4, default 4; must be the same on every rank) cuts a step into pieces
 * whose exchanges overlap the kernels of the other pieces.  While a sharded step runs, the device's persisting-L2 set-aside is switched on
 * for the answer phase only and restored when the call returns (it is a device-wide limit: another context probing a
 * replicated table on the same GPU at that moment runs without it for a few milliseconds).
 * NCCL (libnccl.so.2) is loaded on first use and only when nranks > 1.  kg_comm_init_local puts all ranks into ONE
 * process (peer copies instead of NCCL): tests on a single GPU, or a single process that drives several devices.
 */
#ifndef KMERGUTS_SHARD_H
#define KMERGUTS_SHARD_H

#include "kmerguts.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct kg_comm kg_comm;

#define KG_COMM_ID_BYTES 128
#define KG_SHARD_MAX_RANKS 16

typedef struct kg_shard_stats {        /* of the last kg_batch_run_sharded on this communicator */
    uint64_t keys_sent;                /* lookups routed, all owners (incl. this rank) */
    uint64_t keys_remote;              /* ... of which left this GPU */
    uint64_t keys_received;            /* lookups this rank answered */
    uint64_t replies_sent;             /* hits found in this rank's shard */
    uint64_t replies_received;         /* hits of this rank's sequences */
    uint64_t bytes_sent;               /* over the interconnect (keys + replies, excluding the self segment) */
    float ms_route, ms_keys, ms_answer, ms_replies, ms_merge, ms_total; /* CUDA events; ms_total: host clock around the call */
    int32_t chunks;                    /* pieces the step was cut into (NCCL transport: the exchange of one piece overlaps
                                          the kernels of the others).  1: ms_route .. ms_replies are phase durations;
                                          > 1: the phases overlap and the four values are the times from the start of the
                                          call to the END of that phase for the last piece.  ms_merge is always a duration */
} kg_shard_stats;

/* Rank (0 <= r < nranks) that owns an encoded 8-mer; table builders and loaders partition with it. */
int kg_shard_owner(uint64_t key, int nranks);

int kg_comm_unique_id(uint8_t id[KG_COMM_ID_BYTES]);
int kg_comm_init(kg_context* ctx, int rank, int nranks, const uint8_t id[KG_COMM_ID_BYTES], kg_comm** comm);
/* All ranks in this process: ctxs[r] is the context of rank r (they may share a device, or even be the same context). */
int kg_comm_init_local(kg_context* const* ctxs, int nranks, kg_comm** comms);
void kg_comm_free(kg_comm* comm);
int kg_comm_last_stats(const kg_comm* comm, kg_shard_stats* stats);

/* Shard `rank` of `nranks` of a table: the loaders of kmerguts.h, keeping only the owned signatures.  Reachability under
 * the reference's no-wrap linear probing (KGJ:959-1026) is decided on the WHOLE file image before the split. */
int kg_table_load_sharded(kg_context* ctx, const char* data_dir, int rank, int nranks, kg_table** table);
int kg_table_from_image_sharded(kg_context* ctx, const void* image, size_t nbytes, int rank, int nranks, kg_table** table);
int kg_table_from_device_entries_sharded(kg_context* ctx, const uint64_t* d_keys, const void* d_payload16, size_t n,
                                         int rank, int nranks, kg_table** table);

/* Collective: every rank of the communicator calls it the same number of times, each with its own batch (which may be
 * empty) and the same params.  The result holds the records of THIS rank's sequences. */
int kg_batch_run_sharded(kg_comm* comm, const kg_table* shard, kg_batch* batch, const kg_params* params,
                         kg_result** result);
/* The same for a local communicator group: all ranks in one call. */
int kg_batch_run_sharded_local(kg_comm* const* comms, const kg_table* const* shards, kg_batch* const* batches, int nranks,
                               const kg_params* params, kg_result** results);

#ifdef __cplusplus
}
#endif
#endif /* KMERGUTS_SHARD_H */
