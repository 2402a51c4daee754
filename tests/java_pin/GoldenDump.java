// tests/java_pin/GoldenDump.java -- pins the CPU oracle (oracle/kg_oracle.c) to the REAL KmerGutsJava on any box with a JDK.
//
// This image has no JVM (java, javac, jshell are absent), so the repo's parity chain ends at the C restatement of the
// reference ("parity unpinned").  This class closes the chain in one command on a machine that has a JDK and a checkout
// of rsutormin/KmerGutsJava: it is compiled TOGETHER with the UNMODIFIED lib/src/kmergutsjava/KmerGutsJava.java (same
// package, so that it can reach the public gatherHits / Hit and, by reflection, the private flag fields) and
//
//   kats <kats.txt> <out.txt>                       drives KmerGutsJava.gatherHits (KGJ:457-514 -> processSetOfHits, KGJ:385-455)
//                                                   with the hand-traced vectors of tests/golden/fsm_kats.json and writes the
//                                                   CALL / OTU-COUNTS lines Java prints for them;
//   reports <runs.txt>                              runs KmerGutsJava.main (KGJ:560-654 -> run, KGJ:742-820) once per line of
//                                                   runs.txt: the reference's own two E. coli fixtures against the derived
//                                                   configs[0] table, protein mode and 6-frame mode, under the four flag sets
//                                                   of SURVEY.md 8(d) (the -d runs on a prefix of each fixture: their
//                                                   after-hit / after-call dumps grow quadratically).
//
// tests/java_pin/pin_oracle.sh runs both and compares with what the oracle gives (tests/java_pin/pin_oracle.py compare) and with
// the SHA-256 values committed in tests/golden/c0_report_sha256.json.  Nothing here is used by the product or the tests.
package kmergutsjava;

import java.io.BufferedReader;
import java.io.FileReader;
import java.io.FileWriter;
import java.io.PrintWriter;
import java.lang.reflect.Field;
import java.lang.reflect.Method;
import java.util.ArrayList;
import java.util.List;

public class GoldenDump {
    static void set(KmerGutsJava k, String name, Object v) throws Exception {
        Field f = KmerGutsJava.class.getDeclaredField(name);
        f.setAccessible(true);
        f.set(k, v);
    }

    // kats.txt (written by pin_oracle.py): "KAT name min_hits max_gap min_weighted order nhits" then nhits lines
    // "pos fI oI weightBits avgOffFromEnd" (the weight as the int bits of the float, so that nothing is lost in text)
    static void kats(String in, String out) throws Exception {
        try (BufferedReader br = new BufferedReader(new FileReader(in)); PrintWriter pw = new PrintWriter(new FileWriter(out))) {
            String line;
            while ((line = br.readLine()) != null) {
                if (!line.startsWith("KAT ")) continue;
                String[] h = line.split(" ");
                KmerGutsJava k = new KmerGutsJava();
                set(k, "minHits", Integer.parseInt(h[2]));
                set(k, "maxGap", Integer.parseInt(h[3]));
                set(k, "minWeightedHits", Integer.parseInt(h[4]));
                set(k, "orderConstraint", Integer.parseInt(h[5]) != 0);
                int n = Integer.parseInt(h[6]), maxFI = 0;
                List<KmerGutsJava.Hit> hits = new ArrayList<KmerGutsJava.Hit>();
                for (int i = 0; i < n; i++) {
                    String[] t = br.readLine().trim().split(" ");
                    KmerGutsJava.Hit x = new KmerGutsJava.Hit();
                    x.from0InProt = Integer.parseInt(t[0]);
                    x.fI = Integer.parseInt(t[1]);
                    x.oI = Integer.parseInt(t[2]);
                    x.functionWt = Float.intBitsToFloat(Integer.parseInt(t[3]));
                    x.avgOffFromEnd = Integer.parseInt(t[4]);
                    maxFI = Math.max(maxFI, x.fI);
                    hits.add(x);
                }
                List<String> functions = new ArrayList<String>();
                for (int i = 0; i <= maxFI; i++) functions.add("F" + i);
                List<KmerGutsJava.OtuCount> otu = new ArrayList<KmerGutsJava.OtuCount>();
                pw.println("KAT " + h[1]);
                k.gatherHits(100000, '+', 0, hits, functions, otu, pw);                  // KGJ:457
                Method tab = KmerGutsJava.class.getDeclaredMethod("tabulateOtuDataForContig", String.class, int.class, List.class, PrintWriter.class);
                tab.setAccessible(true);
                tab.invoke(k, "kat", 0, otu, pw);                                         // KGJ:516
            }
        }
    }

    // runs.txt (written by pin_oracle.py): one run per line, tab-separated: the arguments of KmerGutsJava.main
    static void reports(String runs) throws Exception {
        try (BufferedReader br = new BufferedReader(new FileReader(runs))) {
            String line;
            while ((line = br.readLine()) != null) {
                if (line.trim().isEmpty()) continue;
                String[] a = line.split("\t");
                System.err.println("KmerGutsJava.main " + String.join(" ", a));
                KmerGutsJava.main(a);                                                        // KGJ:560
            }
        }
    }

    public static void main(String[] args) throws Exception {
        if (args.length == 3 && args[0].equals("kats")) kats(args[1], args[2]);
        else if (args.length == 2 && args[0].equals("reports")) reports(args[1]);
        else {
            System.err.println("usage: GoldenDump kats <kats.txt> <out.txt> | reports <runs.txt>");
            System.exit(2);
        }
    }
}
