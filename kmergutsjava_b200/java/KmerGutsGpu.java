package kmergutsjava;

import java.lang.foreign.Arena;
import java.lang.foreign.FunctionDescriptor;
import java.lang.foreign.Linker;
import java.lang.foreign.MemoryLayout;
import java.lang.foreign.MemorySegment;
import java.lang.foreign.StructLayout;
import java.lang.foreign.SymbolLookup;
import java.lang.foreign.ValueLayout;
import java.lang.invoke.MethodHandle;
import java.nio.charset.StandardCharsets;
import java.util.ArrayList;
import java.util.List;

import static java.lang.foreign.ValueLayout.ADDRESS;
import static java.lang.foreign.ValueLayout.JAVA_FLOAT;
import static java.lang.foreign.ValueLayout.JAVA_INT;
import static java.lang.foreign.ValueLayout.JAVA_LONG;

/**
 * Panama FFM (JDK 22+) binding of libkmerguts_b200.so (include/kmerguts.h) -- SOURCE ONLY: this image has no JVM, so
 * this file has never been compiled; the same C ABI is exercised by the kmer_guts_b200 command line and by the ctypes
 * tests.  It is what a maintainer drops next to KmerGutsJava.java; INTEGRATION.md shows the 30-line change in run().
 *
 * Replaces, inside KmerGutsJava.run(): prepareQuery (KGJ:1051), createKmerStorage/finalizeSorting (KGJ:822, 846),
 * lookup (KGJ:944) and the gatherHits/processSetOfHits calls (KGJ:457, 385).  FASTA parsing, function.index and the
 * text report stay in Java.
 */
public final class KmerGutsGpu implements AutoCloseable {
    public static final int MODE_DNA = 0, MODE_AA = 1;

    /** struct kg_params */
    static final StructLayout PARAMS = MemoryLayout.structLayout(
            JAVA_INT.withName("min_hits"), JAVA_INT.withName("min_weighted_hits"), JAVA_INT.withName("max_gap"),
            JAVA_INT.withName("order_constraint"), JAVA_INT.withName("emit_hits"));
    /** struct kg_call (32 bytes) */
    static final StructLayout CALL = MemoryLayout.structLayout(
            JAVA_INT.withName("seq"), JAVA_INT.withName("strand_frame"), JAVA_INT.withName("start"), JAVA_INT.withName("end"),
            JAVA_INT.withName("count"), JAVA_INT.withName("fI"), JAVA_FLOAT.withName("weighted"), JAVA_INT.withName("hits_before"));
    /** struct kg_otu (44 bytes) */
    static final StructLayout OTU = MemoryLayout.structLayout(
            JAVA_INT.withName("n"), MemoryLayout.sequenceLayout(5, JAVA_INT).withName("count"),
            MemoryLayout.sequenceLayout(5, JAVA_INT).withName("oI"));
    /** struct kg_hit (28 bytes) */
    static final StructLayout HIT = MemoryLayout.structLayout(
            JAVA_INT.withName("seq"), JAVA_INT.withName("strand_frame"), JAVA_INT.withName("pos"), JAVA_INT.withName("oI"),
            JAVA_INT.withName("avg_off_from_end"), JAVA_INT.withName("fI"), JAVA_FLOAT.withName("function_wt"));

    public record Call(int seq, int strandFrame, int start, int end, int count, int fI, float weighted, int hitsBefore) {}
    public record Otu(int[] count, int[] oI) {}
    public record Result(List<Call> calls, List<Otu> otus, List<KmerGutsJava.Hit>[] hitsBySeqFrame) {}

    private static final Linker LINKER = Linker.nativeLinker();
    private final Arena arena = Arena.ofShared();
    private final SymbolLookup lib;
    private final MethodHandle kgInit, kgShutdown, kgLastError, kgTableLoad, kgTableFree, kgRun, kgResultCalls, kgResultOtus,
            kgResultHits, kgResultFree;
    // hash-sharded table across GPUs (include/kmerguts_shard.h); the reference has no counterpart
    private final MethodHandle kgCommUniqueId, kgCommInit, kgCommFree, kgTableLoadSharded, kgBatchUpload, kgBatchFree, kgBatchRunSharded;
    private final MemorySegment ctx;
    private MemorySegment table = MemorySegment.NULL;
    private MemorySegment comm = MemorySegment.NULL;

    public KmerGutsGpu(String libraryPath, int device) throws Throwable {
        lib = SymbolLookup.libraryLookup(libraryPath, arena);
        kgInit = h("kg_init", FunctionDescriptor.of(JAVA_INT, JAVA_INT, ADDRESS));
        kgShutdown = h("kg_shutdown", FunctionDescriptor.ofVoid(ADDRESS));
        kgLastError = h("kg_last_error", FunctionDescriptor.of(ADDRESS));
        kgTableLoad = h("kg_table_load", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, ADDRESS));
        kgTableFree = h("kg_table_free", FunctionDescriptor.ofVoid(ADDRESS));
        kgRun = h("kg_run", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, JAVA_INT, ADDRESS, ADDRESS, JAVA_LONG, ADDRESS, ADDRESS));
        kgResultCalls = h("kg_result_calls", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, ADDRESS));
        kgResultOtus = h("kg_result_otus", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, ADDRESS));
        kgResultHits = h("kg_result_hits", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, ADDRESS));
        kgResultFree = h("kg_result_free", FunctionDescriptor.ofVoid(ADDRESS));
        kgCommUniqueId = h("kg_comm_unique_id", FunctionDescriptor.of(JAVA_INT, ADDRESS));
        kgCommInit = h("kg_comm_init", FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_INT, JAVA_INT, ADDRESS, ADDRESS));
        kgCommFree = h("kg_comm_free", FunctionDescriptor.ofVoid(ADDRESS));
        kgTableLoadSharded = h("kg_table_load_sharded", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, JAVA_INT, JAVA_INT, ADDRESS));
        kgBatchUpload = h("kg_batch_upload", FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_INT, ADDRESS, ADDRESS, JAVA_LONG, ADDRESS));
        kgBatchFree = h("kg_batch_free", FunctionDescriptor.ofVoid(ADDRESS));
        kgBatchRunSharded = h("kg_batch_run_sharded", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, ADDRESS, ADDRESS, ADDRESS));
        MemorySegment out = arena.allocate(ADDRESS);
        check((int) kgInit.invokeExact(device, out));
        ctx = out.get(ADDRESS, 0);
    }

    private MethodHandle h(String name, FunctionDescriptor fd) {
        return LINKER.downcallHandle(lib.find(name).orElseThrow(() -> new UnsatisfiedLinkError(name)), fd);
    }

    private void check(int rc) throws Throwable {
        if (rc != 0) {
            MemorySegment msg = ((MemorySegment) kgLastError.invokeExact()).reinterpret(1024);
            throw new IllegalStateException("kmerguts error " + rc + ": " + msg.getString(0));
        }
    }

    /** readKmerTableHeader + the table side of lookup: kmer.table.mem_map[.gz] under kmerTableDir (KGJ:749-753, 924-942). */
    public void loadTable(String kmerTableDir) throws Throwable {
        try (Arena a = Arena.ofConfined()) {
            MemorySegment out = a.allocate(ADDRESS);
            check((int) kgTableLoad.invokeExact(ctx, a.allocateFrom(kmerTableDir), out));
            table = out.get(ADDRESS, 0);
        }
    }

    /**
     * prepareQuery + lookup + gatherHits for a batch of FASTA records (the `seq` strings FastaCallback.nextEntry
     * receives, KGJ:780).  aa/minHits/minWeightedHits/maxGap/orderConstraint/debug are KmerGutsJava's instance fields.
     */
    public Result run(List<String> seqs, boolean aa, int minHits, int minWeightedHits, int maxGap, boolean orderConstraint,
                      boolean debug) throws Throwable {
        try (Arena a = Arena.ofConfined()) {
            MemorySegment[] in = pack(a, seqs, minHits, minWeightedHits, maxGap, orderConstraint, debug);
            MemorySegment out = a.allocate(ADDRESS);
            check((int) kgRun.invokeExact(ctx, table, aa ? MODE_AA : MODE_DNA, in[0], in[1], (long) seqs.size(), in[2], out));
            return readAndFree(a, out.get(ADDRESS, 0));
        }
    }

    /** Rank 0 of a multi-GPU job makes the communicator id; the application hands the 128 bytes to every rank. */
    public byte[] commUniqueId() throws Throwable {
        try (Arena a = Arena.ofConfined()) {
            MemorySegment id = a.allocate(128);
            check((int) kgCommUniqueId.invokeExact(id));
            return id.toArray(ValueLayout.JAVA_BYTE);
        }
    }

    /** Collective over the ranks: join the communicator and load this rank's shard of kmer.table.mem_map[.gz]. */
    public void joinSharded(int rank, int nranks, byte[] id, String kmerTableDir) throws Throwable {
        try (Arena a = Arena.ofConfined()) {
            MemorySegment out = a.allocate(ADDRESS);
            MemorySegment idSeg = a.allocate(128);
            MemorySegment.copy(id, 0, idSeg, ValueLayout.JAVA_BYTE, 0, 128);
            check((int) kgCommInit.invokeExact(ctx, rank, nranks, idSeg, out));
            comm = out.get(ADDRESS, 0);
            check((int) kgTableLoadSharded.invokeExact(ctx, a.allocateFrom(kmerTableDir), rank, nranks, out));
            table = out.get(ADDRESS, 0);
        }
    }

    /** run() against the sharded table: collective, every rank calls it the same number of times with its own sequences. */
    public Result runSharded(List<String> seqs, boolean aa, int minHits, int minWeightedHits, int maxGap, boolean orderConstraint,
                             boolean debug) throws Throwable {
        try (Arena a = Arena.ofConfined()) {
            MemorySegment[] in = pack(a, seqs, minHits, minWeightedHits, maxGap, orderConstraint, debug);
            MemorySegment out = a.allocate(ADDRESS);
            check((int) kgBatchUpload.invokeExact(ctx, aa ? MODE_AA : MODE_DNA, in[0], in[1], (long) seqs.size(), out));
            MemorySegment batch = out.get(ADDRESS, 0);
            try {
                check((int) kgBatchRunSharded.invokeExact(comm, table, batch, in[2], out));
                return readAndFree(a, out.get(ADDRESS, 0));
            } finally {
                kgBatchFree.invokeExact(batch);
            }
        }
    }

    /** {sequence bytes, offsets[n+1], kg_params} in native memory. */
    private MemorySegment[] pack(Arena a, List<String> seqs, int minHits, int minWeightedHits, int maxGap, boolean orderConstraint,
                                 boolean debug) {
        long total = 0;
        for (String s : seqs) total += s.length();
        MemorySegment bytes = a.allocate(Math.max(total, 1));
        MemorySegment offs = a.allocate(JAVA_LONG, seqs.size() + 1L);
        long at = 0;
        for (int i = 0; i < seqs.size(); i++) {
            // ISO_8859_1 keeps one byte per Java char, as toAminoAcidOff / dnaChar see them (KGJ:1057, 324)
            byte[] b = seqs.get(i).getBytes(StandardCharsets.ISO_8859_1);
            MemorySegment.copy(b, 0, bytes, ValueLayout.JAVA_BYTE, at, b.length);
            offs.setAtIndex(JAVA_LONG, i, at);
            at += b.length;
        }
        offs.setAtIndex(JAVA_LONG, seqs.size(), at);
        MemorySegment p = a.allocate(PARAMS);
        p.set(JAVA_INT, 0, minHits);
        p.set(JAVA_INT, 4, minWeightedHits);
        p.set(JAVA_INT, 8, maxGap);
        p.set(JAVA_INT, 12, orderConstraint ? 1 : 0);
        p.set(JAVA_INT, 16, debug ? 1 : 0);
        return new MemorySegment[] {bytes, offs, p};
    }

    private Result readAndFree(Arena a, MemorySegment res) throws Throwable {
        try {
            MemorySegment ptr = a.allocate(ADDRESS), cnt = a.allocate(JAVA_LONG);
            check((int) kgResultCalls.invokeExact(res, ptr, cnt));
            long n = cnt.get(JAVA_LONG, 0);
            MemorySegment cs = ptr.get(ADDRESS, 0).reinterpret(n * CALL.byteSize());
            List<Call> calls = new ArrayList<>((int) n);
            for (long i = 0; i < n; i++) {
                long o = i * CALL.byteSize();
                calls.add(new Call(cs.get(JAVA_INT, o), cs.get(JAVA_INT, o + 4), cs.get(JAVA_INT, o + 8), cs.get(JAVA_INT, o + 12),
                        cs.get(JAVA_INT, o + 16), cs.get(JAVA_INT, o + 20), cs.get(JAVA_FLOAT, o + 24), cs.get(JAVA_INT, o + 28)));
            }
            check((int) kgResultOtus.invokeExact(res, ptr, cnt));
            n = cnt.get(JAVA_LONG, 0);
            MemorySegment os = ptr.get(ADDRESS, 0).reinterpret(n * OTU.byteSize());
            List<Otu> otus = new ArrayList<>((int) n);
            for (long i = 0; i < n; i++) {
                long o = i * OTU.byteSize();
                int k = os.get(JAVA_INT, o);
                int[] c = new int[k], oi = new int[k];
                for (int j = 0; j < k; j++) {
                    c[j] = os.get(JAVA_INT, o + 4 + 4L * j);
                    oi[j] = os.get(JAVA_INT, o + 24 + 4L * j);
                }
                otus.add(new Otu(c, oi));
            }
            return new Result(calls, otus, null); // "-d": kg_result_hits is read the same way (HIT layout above)
        } finally {
            kgResultFree.invokeExact(res);
        }
    }

    @Override
    public void close() {
        try {
            if (!comm.equals(MemorySegment.NULL)) kgCommFree.invokeExact(comm);
            if (!table.equals(MemorySegment.NULL)) kgTableFree.invokeExact(table);
            kgShutdown.invokeExact(ctx);
        } catch (Throwable t) {
            throw new RuntimeException(t);
        }
        arena.close();
    }
}
