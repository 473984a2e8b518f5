package jgi;

import java.io.PrintStream;
import java.util.ArrayList;
import java.util.concurrent.atomic.AtomicLongArray;

import stream.Read;

/**
 * GPU-backed loader and matching block of Seal: forwards whole read lists to libbbduk_b200.so
 * (include/seal_b200.h) through jni/SealCuda.c. Written against jgi/Seal.java as it stands in the reference;
 * NOT compiled in this repository's image (no JDK) -- tests/test_java_shim_cpu.py checks the native
 * declarations against the shim and the patch against the members declared here.
 *
 * Seal with gpu=t (java/patches/Seal.diff): LoadThread 0 hands every reference read to addRef in input order
 * (ids stay the reference's: 1, 2, ...), finalizeTable() replaces the tables, ProcessThread.run asks match()
 * once per ListNum and reads sites / assigned per fragment, addScaffoldCounts() folds the per-reference
 * counters in at the end.
 */
public final class SealGPU {

	/*--------------------------------------------------------------*/
	/*----------------        Native Methods        ----------------*/
	/*--------------------------------------------------------------*/

	static native long createNative(int[] cfg17, float clearzoneFraction, float minKmerFraction, int device);
	static native int addRefNative(long handle, byte[] bases);
	static native int finalizeNative(long handle, long[] out3);
	static native int processNative(long handle, byte[] bases, long[] offsets, long nReads, boolean paired, long firstNumericId,
			int idsStride, int[] nAssigned, int[] firstId, int[] nSites, int[] maxHits, int[] ids, long[] stats8);
	static native int scaffoldCountsNative(long handle, long[] reads, long[] bases, long[] frags, long[] ambig);
	static native String lastErrorNative(long handle);
	static native void destroyNative(long handle);

	static{
		System.loadLibrary("bbtoolsjni_b200");//jni/SealCuda.c + jni/BBDukCuda.c linked against libbbduk_b200.so
	}

	/*--------------------------------------------------------------*/
	/*----------------         Construction         ----------------*/
	/*--------------------------------------------------------------*/

	/**
	 * Returns a GPU engine, or null (with a message) when a flag outside the device path is on; Seal then runs as usual.
	 * The arguments are Seal's fields after its constructor's derivations (jgi/Seal.java:486-571); the library repeats the
	 * derivations, which are idempotent on derived values.
	 */
	public static SealGPU createIfServed(int k, boolean rcomp, boolean maskMiddle, int midMaskLen, boolean forbidNs, int hammingDistance,
			int editDistance, int qHammingDistance, int speed, int qSkip, int refSkip, int restrictLeft, int restrictRight, int ambigMode,
			int matchMode, boolean keepPairsTogether, int clearzone, float clearzoneFraction, int minKmerHits, float minKmerFraction,
			boolean preambleIdle, PrintStream outstream){
		if(editDistance>0 || qHammingDistance>0 || hammingDistance>2 || !preambleIdle){
			outstream.println("gpu=t ignored: edist, qhdist, hdist>2, trimming / quality filters, rename, pattern output, pcr and barcodes stay on the CPU path.");
			return null;
		}
		final int[] cfg={k, rcomp ? 1 : 0, maskMiddle ? 1 : 0, midMaskLen, forbidNs ? 1 : 0, hammingDistance, speed, qSkip, refSkip,
				restrictLeft, restrictRight, ambigMode, matchMode, keepPairsTogether ? 1 : 0, clearzone, minKmerHits, 0};
		final long h=createNative(cfg, clearzoneFraction, minKmerFraction, 0);
		if(h==0){
			outstream.println("gpu=t ignored: "+lastErrorNative(0));
			return null;
		}
		return new SealGPU(h, keepPairsTogether);
	}

	private SealGPU(long handle_, boolean keepPairsTogether_){
		handle=handle_;
		keepPairsTogether=keepPairsTogether_;
	}

	/*--------------------------------------------------------------*/
	/*----------------            Loader            ----------------*/
	/*--------------------------------------------------------------*/

	/** One reference read, in input order (replaces LoadThread.addToMap(Read, skip), jgi/Seal.java:1760-1829). */
	public synchronized void addRef(byte[] bases){
		final int rc=addRefNative(handle, bases==null ? new byte[0] : bases);
		if(rc!=0){throw new RuntimeException(lastErrorNative(handle));}
		scaffolds++;
	}

	/** Builds the device table; returns storedKmers ("Added N kmers", jgi/Seal.java:744). */
	public long finalizeTable(){
		final long[] out=new long[3];
		if(finalizeNative(handle, out)!=0){throw new RuntimeException(lastErrorNative(handle));}
		refKmers=out[2];
		return out[0];
	}

	public long refKmers(){return refKmers;}

	/*--------------------------------------------------------------*/
	/*----------------           Matching           ----------------*/
	/*--------------------------------------------------------------*/

	/** Answers of one read list, by index into the list. */
	public static final class Result {
		Result(int n){sites=new int[n]; assigned=new int[n]; removed=new boolean[n];}
		/** finalList sizes (both mates added when pairs are taken apart) */
		public final int[] sites;
		/** references the fragment was assigned to */
		public final int[] assigned;
		/** the length rules of the preamble removed the fragment: it was not matched */
		public final boolean[] removed;
		public long readsMatched, basesMatched, readsUnmatched, basesUnmatched;
	}

	/**
	 * Replaces the block from "Do kmer matching" on (jgi/Seal.java:2186-2276) for a whole list. With every other step of the
	 * preamble idle, a fragment is removed by its lengths alone (:2108-2139); removed fragments are not sent, and because
	 * ambig=random reads Read.numericID (:2403) the list goes out in runs of consecutive ids.
	 */
	public Result match(ArrayList<Read> reads, int minReadLength, int maxReadLength, float minLenFraction, boolean removePairsIfEitherBad){
		final int n=reads.size();
		final Result res=new Result(n);
		int runStart=-1;
		for(int i=0; i<n; i++){
			final Read r1=reads.get(i), r2=r1.mate;
			final int len1=r1.length(), len2=r1.mateLength();
			final int minlen1=(int)Math.max(len1*minLenFraction, minReadLength);
			final int minlen2=(int)Math.max(len2*minLenFraction, minReadLength);
			final boolean bad1=(len1<minlen1 || len1>maxReadLength);
			final boolean bad2=(r2!=null && (len2<minlen2 || len2>maxReadLength));
			res.removed[i]=(removePairsIfEitherBad ? (bad1 || bad2) : (bad1 && (r2==null || bad2)));
			final boolean continues=(runStart>=0 && !res.removed[i] && reads.get(i-1).numericID+1==r1.numericID
					&& (reads.get(i-1).mate==null)==(r2==null));
			if(runStart>=0 && !continues){
				matchRun(reads, runStart, i, res);
				runStart=-1;
			}
			if(!res.removed[i] && runStart<0){runStart=i;}
		}
		if(runStart>=0){matchRun(reads, runStart, n, res);}
		return res;
	}

	private void matchRun(ArrayList<Read> reads, int from, int to, Result res){
		final boolean paired=(reads.get(from).mate!=null);
		final int per=(paired ? 2 : 1), nReads=(to-from)*per;
		final long[] offsets=new long[nReads+1];
		long total=0;
		for(int i=from, j=0; i<to; i++){
			final Read r1=reads.get(i);
			total+=r1.length();
			offsets[++j]=total;
			if(paired){
				total+=r1.mateLength();
				offsets[++j]=total;
			}
		}
		final byte[] bases=new byte[(int)total];
		for(int i=from, j=0; i<to; i++){
			final Read r1=reads.get(i);
			if(r1.bases!=null){System.arraycopy(r1.bases, 0, bases, (int)offsets[j], r1.bases.length);}
			j++;
			if(paired){
				if(r1.mate.bases!=null){System.arraycopy(r1.mate.bases, 0, bases, (int)offsets[j], r1.mate.bases.length);}
				j++;
			}
		}
		final int units=(paired && keepPairsTogether ? nReads/2 : nReads);
		final int[] nAssigned=new int[units], firstId=new int[units], nSites=new int[units], maxHits=new int[units];
		final long[] stats=new long[8];
		final int rc=processNative(handle, bases, offsets, nReads, paired, reads.get(from).numericID, 0, nAssigned, firstId, nSites, maxHits,
				null, stats);
		if(rc!=0){throw new RuntimeException(lastErrorNative(handle));}
		for(int i=from, u=0; i<to; i++){
			if(paired && !keepPairsTogether){//sites and assigned of both mates (jgi/Seal.java:2270-2272)
				res.sites[i]=nSites[u]+nSites[u+1];
				res.assigned[i]=nAssigned[u]+nAssigned[u+1];
				u+=2;
			}else{
				res.sites[i]=nSites[u];
				res.assigned[i]=nAssigned[u];
				u++;
			}
		}
		res.readsMatched+=stats[2];
		res.basesMatched+=stats[3];
		res.readsUnmatched+=stats[4];
		res.basesUnmatched+=stats[5];
	}

	/** Folds the device's per-reference totals into Seal's arrays (jgi/Seal.java:2431-2441), once, after the ProcessThreads. */
	public void addScaffoldCounts(AtomicLongArray reads, AtomicLongArray bases, AtomicLongArray frags, AtomicLongArray ambig){
		if(reads==null || bases==null || frags==null || ambig==null){return;}
		final int n=scaffolds+1;
		final long[] r=new long[n], b=new long[n], f=new long[n], a=new long[n];
		if(scaffoldCountsNative(handle, r, b, f, a)!=0){throw new RuntimeException(lastErrorNative(handle));}
		for(int i=0; i<n && i<reads.length(); i++){
			reads.addAndGet(i, r[i]);
			bases.addAndGet(i, b[i]);
			frags.addAndGet(i, f[i]);
			ambig.addAndGet(i, a[i]);
		}
	}

	public void close(){
		if(handle!=0){destroyNative(handle);}
		handle=0;
	}

	/*--------------------------------------------------------------*/
	/*----------------            Fields            ----------------*/
	/*--------------------------------------------------------------*/

	private long handle;
	private final boolean keepPairsTogether;
	private int scaffolds=0;
	private long refKmers=0;

}
