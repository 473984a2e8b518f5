package kmer;

import java.util.ArrayList;

import stream.Read;

/**
 * GPU-backed counting table of KmerCountExact: forwards whole read lists to libbbduk_b200.so
 * (include/kcount_b200.h) through jni/KCountCuda.c. Written against kmer/KmerTableSet.java as it stands in the
 * reference; NOT compiled in this repository's image (no JDK) -- tests/test_java_shim_cpu.py checks the native
 * declarations against the shim.
 *
 * KmerTableSet.LoadThread (kmer/KmerTableSet.java:489-592) aggregates the reads of its ListNums and calls
 * addReads instead of addKmersToTable (:652-716); AbstractKmerTableSet.fillHistogram (:370) becomes khist();
 * "Unique Kmers" (jgi/KmerCountExact.java:355-367) comes from stats()[3].
 */
public final class KmerTableSetGPU {

	static native long createNative(int k, boolean rcomp, long initialKeys);
	static native int addReadsNative(long handle, byte[] bases, long[] offsets, long nReads);
	static native int statsNative(long handle, long[] stats4);
	static native int khistNative(long handle, int histMax, long[] hist);
	static native String lastErrorNative(long handle);
	static native void destroyNative(long handle);

	static{
		System.loadLibrary("bbtoolsjni_b200");
	}

	public KmerTableSetGPU(int k, boolean rcomp, long initialKeys){
		handle=createNative(k, rcomp, initialKeys);
		if(handle==0){throw new RuntimeException(lastErrorNative(0));}
	}

	/** Counts every k-mer of the list's reads and mates (replaces addKmersToTable per read, kmer/KmerTableSet.java:652-716). */
	public synchronized void addReads(ArrayList<Read> reads){
		int n=0;
		long total=0;
		for(Read r : reads){
			n+=1+(r.mate==null ? 0 : 1);
			total+=r.length()+r.mateLength();
		}
		final long[] offsets=new long[n+1];
		final byte[] bases=new byte[(int)total];
		int j=0;
		long at=0;
		for(Read r1 : reads){
			for(Read r=r1; r!=null; r=(r==r1 ? r1.mate : null)){
				if(r.bases!=null){System.arraycopy(r.bases, 0, bases, (int)at, r.bases.length);}
				at+=r.length();
				offsets[++j]=at;
			}
		}
		if(addReadsNative(handle, bases, offsets, n)!=0){throw new RuntimeException(lastErrorNative(handle));}
	}

	/** {readsIn, basesIn, kmersIn, unique k-mers} */
	public long[] stats(){
		final long[] v=new long[4];
		if(statsNative(handle, v)!=0){throw new RuntimeException(lastErrorNative(handle));}
		return v;
	}

	/** hist[min(count, histMax)]++ over all keys (kmer/HashArray.java:577-588) */
	public long[] khist(int histMax){
		final long[] hist=new long[histMax+1];
		if(khistNative(handle, histMax, hist)!=0){throw new RuntimeException(lastErrorNative(handle));}
		return hist;
	}

	public void close(){
		if(handle!=0){destroyNative(handle);}
		handle=0;
	}

	private long handle;

}
