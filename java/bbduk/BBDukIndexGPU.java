package bbduk;

import java.util.ArrayList;

import shared.Shared;
import stream.Read;

/**
 * Fourth BBDukIndex implementation (next to BBDukIndexMod / Mask / Mask2): the reference k-mer table
 * lives in GPU memory and whole batches of reads are answered by one native call.
 * NOT compiled in this repository (no JDK in the image); it shows the binding a maintainer adds.
 *
 * Selected in BBDukLoader's constructor (bbduk/BBDukLoader.java:74-75) by one new flag:
 *   index = p.gpu ? new BBDukIndexGPU(p) : (p.WAYS==7 ? new BBDukIndexMod(p) : ...);
 * and used by BBDukProcessorS.processList (bbduk/BBDukProcessorS.java:768) BEFORE the per-read loop:
 * the processor aggregates lists until >= MIN_BATCH reads, calls processBatch once, then walks the reads
 * applying (lo, hi, flags) exactly where ktrim()/kmask()/countSetKmers() used to be called
 * (bbduk/BBDukProcessorS.java:947-1093). The per-k-mer getValue() stays only for dump/verbose.
 */
public final class BBDukIndexGPU extends BBDukIndex {

	static { Shared.loadJNI("bbdukcuda"); } // shared/Shared.java:731-779: searches <classpath>/../jni too

	private long handle;
	public static final int MIN_BATCH=1<<20;

	public BBDukIndexGPU(BBDukParser p){
		super(p);
		handle=createNative(marshal(p));
		if(handle==0){throw new RuntimeException("bbduk_b200_create failed: "+lastErrorNative(0));}
	}

	/** bbduk_cfg in declaration order (include/bbduk_b200.h); floats as raw int bits */
	private static int[] marshal(BBDukParser p){
		return new int[] {0 /*struct_size, set natively*/, 1 /*BBDUK_GEN_S*/, p.kbig>p.k ? p.kbig : p.k, p.mink,
			p.useShortKmers ? 1 : 0, p.hammingDistance, p.hammingDistance2, p.editDistance, p.editDistance2,
			p.qHammingDistance, p.qHammingDistance2, p.rcomp ? 1 : 0, p.maskMiddle ? 1 : 0, p.midMaskLen,
			p.forbidNs ? 1 : 0, p.ktrimLeft ? 1 : 0, p.ktrimRight ? 1 : 0, p.ktrimN ? 1 : 0, p.ksplit ? 1 : 0,
			p.ktrimExclusive ? 1 : 0, p.trimPad, p.restrictLeft, p.restrictRight, p.skipR1 ? 1 : 0, p.skipR2 ? 1 : 0,
			p.qSkip, p.speed, p.minSkip, p.maxSkip, p.maxBadKmers0, Float.floatToRawIntBits(p.minKmerFraction),
			Float.floatToRawIntBits(p.minCoveredFraction), p.findBestMatch ? 1 : 0, p.kmaskFullyCovered ? 1 : 0,
			p.kmaskLowercase ? 1 : 0, p.trimSymbol, p.minReadLength, Float.floatToRawIntBits(p.minLenFraction),
			p.removePairsIfEitherBad ? 0 : 1, p.trimPairsEvenly ? 1 : 0, p.trimFailuresTo1bp ? 1 : 0, -1, 0};
	}

	/** Called by BBDukLoader for every list of scaffolds, in file order (ids continue across calls). */
	public void addScaffolds(ArrayList<Read> scafs){
		int total=0; for(Read r : scafs){total+=r.length();}
		byte[] bases=new byte[total]; long[] off=new long[scafs.size()+1];
		int pos=0, i=0;
		for(Read r : scafs){System.arraycopy(r.bases, 0, bases, pos, r.length()); pos+=r.length(); off[++i]=pos;}
		if(addRefNative(handle, bases, off, scafs.size())!=0){throw new RuntimeException(lastErrorNative(handle));}
	}

	@Override public void setKmersLoaded(){storedKmers=finalizeNative(handle);}

	/** One aggregated batch; mates adjacent. Arrays are reused by the caller between batches. */
	public boolean processBatch(byte[] bases, long[] offsets, long nReads, boolean paired,
			int[] id0, int[] lo, int[] hi, byte[] flags, int[] count, long[] stats8){
		return processNative(handle, bases, offsets, nReads, paired, id0, lo, hi, flags, count, stats8)==0;
	}

	/** Trim by overlap for the batch processBatch() just answered (replaces bbduk/BBDukProcessorS.java:1096-1143):
	 * hi[] / flags[] are updated in place; stats2 += {readsTrimmedByOverlap, basesTrimmedByOverlap}. */
	public boolean tboBatch(boolean strictOverlap, int minOverlap0, int minOverlap, int minInsert0, int minInsert, float meeFilter,
			byte[] bases, byte[] quals, long[] offsets, long nReads, int[] lo, int[] hi, byte[] flags, int[] insert, long[] stats2){
		final int[] cfg={strictOverlap ? 1 : 0, minOverlap0, minOverlap, minInsert0, minInsert, 33};
		return tboNative(handle, cfg, meeFilter, bases, quals, offsets, nReads, lo, hi, flags, insert, stats2)==0;
	}

	/** Poly-X trimming, quality trimming and minlen / maxlen / mbq / maxns for the batch processBatch() (and tboBatch())
	 * answered (replaces jgi/BBDuk.java:2954-3052, :3074-3170): lo[] / hi[] / flags[] are updated in place; stats8 +=
	 * {readsQTrimmed, basesQTrimmed, readsQFiltered, basesQFiltered, readsNFiltered, basesNFiltered, readsPolyTrimmed,
	 * basesPolyTrimmed}. quals = Read.quality, flattened. poly = {trimPolyA, trimPolyGLeft, trimPolyGRight, filterPolyG,
	 * trimPolyCLeft, trimPolyCRight, filterPolyC, maxNonPoly}. */
	public boolean qtrimBatch(boolean qtrimLeft, boolean qtrimRight, float trimq, int minBaseQuality, int maxNs, int maxReadLength,
			int[] poly, byte[] bases, byte[] quals, long[] offsets, long nReads, boolean paired, int[] lo, int[] hi, byte[] flags,
			long[] stats8){
		final int[] cfg={qtrimLeft ? 1 : 0, qtrimRight ? 1 : 0, minBaseQuality, maxNs, maxReadLength, 0,
				poly[0], poly[1], poly[2], poly[3], poly[4], poly[5], poly[6], poly[7]};
		return qtrimNative(handle, cfg, trimq, bases, quals, offsets, nReads, paired, lo, hi, flags, stats8)==0;
	}

	/** Low-entropy read filter for the batch the earlier calls answered (replaces jgi/BBDuk.java:3175-3186): flags[] are
	 * updated in place; stats2 += {readsEFiltered, basesEFiltered}. */
	public boolean entropyBatch(float cutoff, int entropyK, int entropyWindow, boolean highPass, byte[] bases, long[] offsets,
			long nReads, boolean paired, int[] lo, int[] hi, byte[] flags, long[] stats2){
		final int[] cfg={entropyK, entropyWindow, highPass ? 1 : 0};
		return entropyNative(handle, cfg, cutoff, bases, offsets, nReads, paired, lo, hi, flags, stats2)==0;
	}

	@Override public int getValue(long kmer, long rkmer, long lengthMask, int qPos, int len, int qHDist){
		throw new UnsupportedOperationException("per-k-mer queries are served in batches by processBatch()");
	}

	@Override public void cleanup(){if(handle!=0){destroyNative(handle); handle=0;}}

	private static native long createNative(int[] cfg);
	private static native int addRefNative(long h, byte[] bases, long[] offsets, int nSeqs);
	private static native long finalizeNative(long h);
	private static native int processNative(long h, byte[] bases, long[] offsets, long nReads, boolean paired,
			int[] id0, int[] lo, int[] hi, byte[] flags, int[] count, long[] stats8);
	private static native int tboNative(long h, int[] cfg, float meeFilter, byte[] bases, byte[] quals, long[] offsets, long nReads,
			int[] lo, int[] hi, byte[] flags, int[] insert, long[] stats2);
	private static native int qtrimNative(long h, int[] cfg, float trimq, byte[] bases, byte[] quals, long[] offsets, long nReads,
			boolean paired, int[] lo, int[] hi, byte[] flags, long[] stats8);
	private static native int entropyNative(long h, int[] cfg, float cutoff, byte[] bases, long[] offsets, long nReads, boolean paired,
			int[] lo, int[] hi, byte[] flags, long[] stats2);
	/** k-mer block + tbo + poly-X / quality trimming / filters + entropy filter in one native call (one upload of the batch);
	 * steps = {doTbo, doQtrim, doEntropy}, floats = {meeFilter, trimq, entropyCutoff}, stats28 as documented in jni/BBDukCuda.c. */
	static native int processChainNative(long h, int[] steps, int[] tboCfg, int[] qCfg, int[] eCfg, float[] floats, byte[] bases,
			byte[] quals, long[] offsets, long nReads, boolean paired, int[] id0, int[] lo, int[] hi, byte[] flags, long[] stats28);
	private static native int scaffoldCountsNative(long h, long[] reads, long[] bases);
	private static native String lastErrorNative(long h);
	private static native void destroyNative(long h);
}
