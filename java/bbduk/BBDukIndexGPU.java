package bbduk;

import java.util.ArrayList;

import aligner.SideChannel4;
import fileIO.ByteStreamWriter;
import fileIO.ReadWrite;
import fileIO.TextStreamWriter;
import shared.Shared;
import shared.Tools;
import stream.Read;

/**
 * Fourth BBDukIndex implementation (next to BBDukIndexMod / Mask / Mask2): the reference k-mer table
 * lives in GPU memory (one replica per GPU of the box) and whole batches of reads are answered by one
 * native call into libbbduk_b200.so through jni/BBDukCuda.c.
 *
 * Not compiled in this repository (there is no JDK in the build image). It is written against
 * bbduk/BBDukIndex.java as it stands: every abstract method of that class is implemented below
 * (tests/test_java_shim_cpu.py checks the list), the constructor relies on the implicit superclass constructor, and the three
 * edits to the reference that select and call it are java/patches/*.diff.
 *
 * Selected in BBDukLoader's constructor (bbduk/BBDukLoader.java:74-75) by the new flag gpu=t; filled by
 * BBDukLoader.LoadThread (one thread: ways() is 1) handing over whole scaffold lists instead of single
 * k-mers; used by BBDukProcessorS.processList (bbduk/BBDukProcessorS.java:768) through BBDukGpuBatch BEFORE
 * the per-read loop. The per-k-mer getValue()/addToMap() stay unimplemented on purpose: the named loop has
 * no CPU path.
 */
public final class BBDukIndexGPU extends BBDukIndex {

	static { Shared.loadJNI("bbdukcuda"); } // shared/Shared.java:731-779: searches <classpath>/../jni too

	/*--------------------------------------------------------------*/
	/*----------------        Initialization        ----------------*/
	/*--------------------------------------------------------------*/

	public BBDukIndexGPU(BBDukParser p){
		scaffoldNames.add(""); //Necessary so that the first real scaffold gets an id of 1, not zero
		scaffoldLengths.add(0);

		refNames=p.refNames;
		altRefNames=p.altRefNames;
		ref=p.ref;
		altref=p.altref;
		literal=p.literal;
		samref=p.samref;
		outrefstats=p.outrefstats;
		printNonZeroOnly=p.printNonZeroOnly;
		overwrite=BBDukParser.overwrite;
		k=p.k;
		unsupportedAssorted=(p.varFile!=null || p.vcfFile!=null || p.filterVars || p.align);
		refScafCounts=new int[refNames.size()];

		final int[] cfg=marshal(p);
		handle=createNative(cfg);
		if(handle==0){throw new RuntimeException("bbduk_b200_create failed: "+lastErrorNative(0));}

		//The library derives its constants again from what marshal() sent; they must agree with the parser's
		final long[] v=new long[16];
		if(describeNative(cfg, v)!=0){throw new RuntimeException("bbduk_b200_describe_cfg failed: "+lastErrorNative(0));}
		assert(v[0]==p.k && v[2]==p.mink && (v[3]!=0)==p.useShortKmers) : "k/mink/useShortKmers differ: "+v[0]+", "+v[2]+", "+v[3];
		assert(v[7]==p.minlen2) : "minlen2 differs: native "+v[7]+", parser "+p.minlen2;
		assert((v[9]!=0)==p.forbidNs) : "forbidNs differs";
		assert(v[12]==p.middleMask) : "middleMask differs: native "+v[12]+", parser "+p.middleMask;
		assert(v[13]==p.mask) : "mask differs";

		gpus=(p.gpuDevices==null || p.gpuDevices.length<1 ? new int[] {0} : p.gpuDevices);
	}

	/**
	 * bbduk_cfg in declaration order (include/bbduk_b200.h); floats as raw int bits. BBDukParser's fields are
	 * already DERIVED when this runs (hammingDistance maxed with editDistance, forbidNs or-ed with hdist<1,
	 * maskMiddle / midMaskLen cleared by useShortKmers or kbig>k). The library derives again; that is idempotent
	 * for everything except minlen2, which the parser computes BEFORE it clears maskMiddle
	 * (bbduk/BBDukParser.java:276 vs :290-294) -- so minlen2 travels explicitly in cfg.minlen2.
	 */
	private static int[] marshal(BBDukParser p){
		return new int[] {0 /*struct_size, set natively*/, 1 /*BBDUK_GEN_S*/, p.kbig>p.k ? p.kbig : p.k, p.mink,
			p.useShortKmers ? 1 : 0, p.hammingDistance, p.hammingDistance2, p.editDistance, p.editDistance2,
			p.qHammingDistance, p.qHammingDistance2, p.rcomp ? 1 : 0, p.maskMiddle ? 1 : 0, p.midMaskLen,
			p.forbidNs ? 1 : 0, p.ktrimLeft ? 1 : 0, p.ktrimRight ? 1 : 0, p.ktrimN ? 1 : 0, p.ksplit ? 1 : 0,
			p.ktrimExclusive ? 1 : 0, p.trimPad, p.restrictLeft, p.restrictRight, p.skipR1 ? 1 : 0, p.skipR2 ? 1 : 0,
			p.qSkip, p.speed, p.minSkip, p.maxSkip, p.maxBadKmers0, Float.floatToRawIntBits(p.minKmerFraction),
			Float.floatToRawIntBits(p.minCoveredFraction), p.findBestMatch ? 1 : 0, p.kmaskFullyCovered ? 1 : 0,
			p.kmaskLowercase ? 1 : 0, p.trimSymbol, p.minReadLength, Float.floatToRawIntBits(p.minLenFraction),
			p.removePairsIfEitherBad ? 0 : 1, p.trimPairsEvenly ? 1 : 0, p.trimFailuresTo1bp ? 1 : 0,
			(p.gpuDevices==null || p.gpuDevices.length<1) ? 0 : p.gpuDevices[0] /*device*/, 0 /*table_load_pct*/,
			p.minlen2 /*cfg.minlen2*/};
	}

	/*--------------------------------------------------------------*/
	/*----------------      Abstract Method Impls   ----------------*/
	/*--------------------------------------------------------------*/

	@Override
	boolean loaded(){return assortedLoaded && kmersLoaded;}

	/** samref / variant loading is host-side bookkeeping and not part of the device path: refuse instead of ignoring it. */
	@Override
	synchronized void loadAssorted(String in1_for_header){
		assert(!assortedLoaded);
		if(samref!=null || unsupportedAssorted){
			throw new UnsupportedOperationException("gpu=t does not support samref= / var= / vcf= / filtervars / align=");
		}
		assortedLoaded=true;
	}

	@Override
	void cleanup(){
		if(BBDukParser.RELEASE_TABLES){
			unloadKmers();
			unloadScaffolds();
		}
	}

	/** Frees the device tables of all replicas. */
	@Override
	synchronized void unloadKmers(){
		if(replicas!=null){
			for(long h : replicas){if(h!=0 && h!=handle){destroyNative(h);}}
			replicas=null;
		}
		if(handle!=0){destroyNative(handle); handle=0;}
	}

	@Override
	void unloadScaffolds(){
		if(scaffoldNames!=null && !scaffoldNames.isEmpty()){
			scaffoldNames.clear();
			scaffoldNames.trimToSize();
		}
		scaffoldReadCounts=null;
		scaffoldBaseCounts=null;
		scaffoldLengths=null;
	}

	/** One loader thread: the device expands every neighbourhood itself, there is nothing to split into ways. */
	@Override
	int ways(){return 1;}

	@Override
	long addToMap(long kmer, long rkmer, int k, long extraBase, int id, long kmask,
			int hammingDistance, int editDistance, int tnum){
		throw new UnsupportedOperationException("gpu=t loads whole scaffolds (addScaffolds), not single kmers");
	}

	@Override
	long addToMapRightShift(long kmer, long rkmer, int id, int tnum){
		throw new UnsupportedOperationException("gpu=t loads whole scaffolds (addScaffolds), not single kmers");
	}

	@Override
	long addToMapLeftShift(long kmer, long rkmer, long extraBase, int id, int tnum){
		throw new UnsupportedOperationException("gpu=t loads whole scaffolds (addScaffolds), not single kmers");
	}

	/** Nothing to rebalance: the device array is sized once from the neighbourhood bound and shrunk after the build. */
	@Override
	void rebalance(int tnum){}

	/** Dumps the stored keys as the other indices do (kmer.AbstractKmerTable.dumpKmersAsBytes: "kmer\tvalue" lines). */
	@Override
	void dump(ByteStreamWriter bsw, int minValue, int maxValue){
		final long n=storedKmers;
		if(n<1){return;}
		assert(n<Shared.MAX_ARRAY_LEN) : "Too many kmers to dump through one array: "+n;
		final long[] keys=new long[(int)n];
		final int[] ids=new int[(int)n];
		final long got=dumpNative(handle, keys, ids);
		if(got<0){throw new RuntimeException("bbduk_b200 table dump failed: "+lastErrorNative(handle));}
		for(int i=0; i<got; i++){
			final int v=ids[i];
			if(v>=minValue && v<=maxValue){
				//the key carries its length marker (bbduk/BBDukIndexMask2.java:533-545): strip it to print the bases
				final long key=keys[i];
				final int len=(63-Long.numberOfLeadingZeros(key))/2;
				final long kmer=key&~(1L<<(2*len));
				bsw.print(dna.AminoAcid.kmerToString(kmer, len)).tab().print(v).nl();
			}
		}
	}

	@Override
	synchronized void setKmersLoaded(){
		assert(!kmersLoaded);
		kmersLoaded=true;
	}

	/** Fills the scaffold names array with reference names (as bbduk/BBDukIndexMask2.java:243-254). */
	@Override
	void toRefNames(){
		final int numRefs=refNames.size();
		for(int r=0, s=1; r<numRefs; r++){
			final int scafs=refScafCounts[r];
			final int lim=s+scafs;
			final String name=ReadWrite.stripToCore(refNames.get(r));
			while(s<lim){
				scaffoldNames.set(s, name);
				s++;
			}
		}
	}

	/** align= needs the host-side SideChannel4 and is not offered with gpu=t (BBDukParser rejects the combination). */
	@Override
	SideChannel4 sidechannel(){return null;}

	@Override
	int getValue(long kmer, long rkmer, long lengthMask, int qPos, int len, int qHDist){
		throw new UnsupportedOperationException("gpu=t answers whole batches (BBDukGpuBatch), not single kmers");
	}

	/** Write statistics on a per-reference basis (as bbduk/BBDukIndexMask2.java:191-240). */
	@Override
	void writeRefStats(String in1, String in2, long readsIn){
		if(outrefstats==null){return;}
		final TextStreamWriter tsw=new TextStreamWriter(outrefstats, overwrite, false, false);
		tsw.start();

		long mapped=0;
		for(int i=0; i<scaffoldReadCounts.length(); i++){mapped+=scaffoldReadCounts.get(i);}

		final int numRefs=refNames.size();
		final long[] refReadCounts=new long[numRefs];
		final long[] refBaseCounts=new long[numRefs];
		final long[] refLengths=new long[numRefs];
		for(int r=0, s=1; r<numRefs; r++){
			final int lim=s+refScafCounts[r];
			while(s<lim){
				refReadCounts[r]+=scaffoldReadCounts.get(s);
				refBaseCounts[r]+=scaffoldBaseCounts.get(s);
				refLengths[r]+=scaffoldLengths.get(s);
				s++;
			}
		}

		tsw.print("#File\t"+in1+(in2==null ? "" : "\t"+in2)+"\n");
		tsw.print(Tools.format("#Reads\t%d\n",readsIn));
		tsw.print(Tools.format("#Mapped\t%d\n",mapped));
		tsw.print(Tools.format("#References\t%d\n",Tools.max(0, refNames.size())));
		tsw.print("#Name\tLength\tScaffolds\tBases\tCoverage\tReads\tRPKM\n");

		final float mult=1000000000f/Tools.max(1, mapped);
		for(int i=0; i<refNames.size(); i++){
			final long reads=refReadCounts[i];
			final long bases=refBaseCounts[i];
			final long len=refLengths[i];
			final int scafs=refScafCounts[i];
			final String name=ReadWrite.stripToCore(refNames.get(i));
			final double invlen=1.0/Tools.max(1, len);
			final double mult2=mult*invlen;
			if(reads>0 || !printNonZeroOnly){
				tsw.print(Tools.format("%s\t%d\t%d\t%d\t%.4f\t%d\t%.4f\n",name,len,scafs,bases,bases*invlen,reads,reads*mult2));
			}
		}
		tsw.poisonAndWait();
	}

	/*--------------------------------------------------------------*/
	/*----------------      Loading (BBDukLoader)   ----------------*/
	/*--------------------------------------------------------------*/

	/**
	 * Called by BBDukLoader.LoadThread for every list of scaffolds, in file order (scaffold ids continue across
	 * calls, first id 1; the loader numbered them the same way, bbduk/BBDukLoader.java:219-233).
	 * @return {reads, bases} seen, for the loader's refReadsT / refBasesT
	 */
	long[] addScaffolds(ArrayList<Read> scafs){
		long total=0;
		int n=0;
		for(Read r : scafs){
			for(Read x=r; x!=null; x=(x==r ? r.mate : null)){
				if(x.bases!=null){total+=x.length(); n++;}
			}
		}
		assert(total<Shared.MAX_ARRAY_LEN) : "reference list too large for one call: "+total;
		final byte[] bases=new byte[(int)total];
		final long[] off=new long[n+1];
		int pos=0, i=0;
		for(Read r : scafs){
			//a reference "pair" shares one scaffold id in the loader; the device numbers sequences 1..n, so mates are
			//not supported for references (FASTQ.FORCE_INTERLEAVED is off while loading, bbduk/BBDukLoader.java:118-119)
			assert(r.mate==null) : "paired reference sequences are not supported with gpu=t";
			if(r.bases==null){continue;}
			System.arraycopy(r.bases, 0, bases, pos, r.length());
			pos+=r.length();
			off[++i]=pos;
		}
		if(addRefNative(handle, bases, off, n)!=0){throw new RuntimeException(lastErrorNative(handle));}
		return new long[] {n, total};
	}

	/**
	 * Builds the device table (neighbourhoods, short kmers), replicates it to the other GPUs of gpus= with one NCCL
	 * broadcast per blob inside the library, and returns "Added N kmers" (bbduk/BBDukLoader.java:343).
	 */
	long finalizeTable(){
		final long[] out=new long[2];
		if(finalizeNative(handle, out)!=0){throw new RuntimeException(lastErrorNative(handle));}
		refKmersNative=out[1];
		replicas=new long[gpus.length];
		replicas[0]=handle;
		if(gpus.length>1){
			final int[] others=new int[gpus.length-1];
			System.arraycopy(gpus, 1, others, 0, others.length);
			final long[] hs=new long[others.length];
			if(replicateNative(handle, others, hs)!=0){throw new RuntimeException(lastErrorNative(handle));}
			System.arraycopy(hs, 0, replicas, 1, hs.length);
		}
		return out[0];
	}

	/**
	 * Adds the per-scaffold hit counters the replicas accumulated on the device (replaces the scaffoldReadCountsT /
	 * scaffoldBaseCountsT merge of BBDukProcessorS.add) to scaffoldReadCounts / scaffoldBaseCounts and resets nothing:
	 * call it once, after the last ProcessThread has finished (bbduk/BBDukS.java:337).
	 */
	synchronized void fetchScaffoldCounts(){
		if(replicas==null || scaffoldReadCounts==null){return;}
		final int n=scaffoldReadCounts.length();
		final long[] reads=new long[n], bases=new long[n];
		for(long h : replicas){
			if(scaffoldCountsNative(h, reads, bases)!=0){throw new RuntimeException(lastErrorNative(h));}
			for(int i=0; i<n; i++){
				if(reads[i]!=0){scaffoldReadCounts.addAndGet(i, reads[i]);}
				if(bases[i]!=0){scaffoldBaseCounts.addAndGet(i, bases[i]);}
			}
		}
	}

	/** refKmers as the LoadThreads would have counted them (bbduk/BBDukLoader.java:461). */
	long refKmersSeen(){return refKmersNative;}

	/*--------------------------------------------------------------*/
	/*----------------   Processing (BBDukGpuBatch)  ----------------*/
	/*--------------------------------------------------------------*/

	/** The replica a ProcessThread should use: threads are spread over the GPUs round robin. */
	long handleFor(int tnum){return replicas[tnum%replicas.length];}
	int replicaCount(){return replicas==null ? 0 : replicas.length;}

	/** One aggregated batch; mates adjacent (2i, 2i+1). The arrays belong to the caller and are reused between batches.
	 * maskBits / maskOff are used in kmask mode only (may be null otherwise). */
	boolean processBatch(long h, byte[] bases, long[] offsets, long nReads, boolean paired,
			int[] id0, int[] id0b, int[] lo, int[] hi, byte[] flags, int[] count, int[] maskBits, long[] maskOff, long[] stats8){
		return processNative(h, bases, offsets, nReads, paired, id0, id0b, lo, hi, flags, count, maskBits, maskOff, stats8)==0;
	}

	/** Trim by overlap for the batch processBatch() just answered (replaces bbduk/BBDukProcessorS.java:1096-1143):
	 * hi[] / flags[] are updated in place; stats2 += {readsTrimmedByOverlap, basesTrimmedByOverlap}. */
	boolean tboBatch(long h, boolean strictOverlap, int minOverlap0, int minOverlap, int minInsert0, int minInsert, float meeFilter,
			byte[] bases, byte[] quals, long[] offsets, long nReads, int[] lo, int[] hi, byte[] flags, int[] insert, long[] stats2){
		final int[] cfg={strictOverlap ? 1 : 0, minOverlap0, minOverlap, minInsert0, minInsert, 33};
		return tboNative(h, cfg, meeFilter, bases, quals, offsets, nReads, lo, hi, flags, insert, stats2)==0;
	}

	/** Poly-X trimming, quality trimming and minlen / maxlen / mbq / maxns for the batch processBatch() (and tboBatch())
	 * answered (replaces jgi/BBDuk.java:2954-3052, :3074-3170): lo[] / hi[] / flags[] are updated in place; stats8 +=
	 * {readsQTrimmed, basesQTrimmed, readsQFiltered, basesQFiltered, readsNFiltered, basesNFiltered, readsPolyTrimmed,
	 * basesPolyTrimmed}. quals = Read.quality, flattened. poly = {trimPolyA, trimPolyGLeft, trimPolyGRight, filterPolyG,
	 * trimPolyCLeft, trimPolyCRight, filterPolyC, maxNonPoly}. maxNRate (>= 1 = off), minConsecutiveBases and minBaseFrequency
	 * (0 = off) are the filters of jgi/BBDuk.java:3138-3159; the two rates travel as raw float bits in the int vector. */
	boolean qtrimBatch(long h, boolean qtrimLeft, boolean qtrimRight, float trimq, int minBaseQuality, int maxNs, int maxReadLength,
			int[] poly, float maxNRate, int minConsecutiveBases, float minBaseFrequency, byte[] bases, byte[] quals, long[] offsets,
			long nReads, boolean paired, int[] lo, int[] hi, byte[] flags, long[] stats8){
		final int[] cfg={qtrimLeft ? 1 : 0, qtrimRight ? 1 : 0, minBaseQuality, maxNs, maxReadLength, 0,
				poly[0], poly[1], poly[2], poly[3], poly[4], poly[5], poly[6], poly[7],
				minConsecutiveBases, Float.floatToRawIntBits(maxNRate), Float.floatToRawIntBits(minBaseFrequency)};
		return qtrimNative(h, cfg, trimq, bases, quals, offsets, nReads, paired, lo, hi, flags, stats8)==0;
	}

	/** Low-entropy read filter for the batch the earlier calls answered (replaces jgi/BBDuk.java:3175-3186): flags[] are
	 * updated in place; stats2 += {readsEFiltered, basesEFiltered}. */
	boolean entropyBatch(long h, float cutoff, int entropyK, int entropyWindow, boolean highPass, byte[] bases, long[] offsets,
			long nReads, boolean paired, int[] lo, int[] hi, byte[] flags, long[] stats2){
		final int[] cfg={entropyK, entropyWindow, highPass ? 1 : 0};
		return entropyNative(h, cfg, cutoff, bases, offsets, nReads, paired, lo, hi, flags, stats2)==0;
	}

	String lastError(long h){return lastErrorNative(h);}

	/*--------------------------------------------------------------*/
	/*----------------            Fields            ----------------*/
	/*--------------------------------------------------------------*/

	/** Native handle of the table that was built (device gpus[0]); 0 after unloadKmers(). */
	private long handle;
	/** One handle per GPU of gpus=, replicas[0]==handle. */
	private long[] replicas;
	/** CUDA device ordinals (gpus=0,1,2,...; default {0}). */
	private final int[] gpus;
	private long refKmersNative;
	private boolean assortedLoaded=false;
	private boolean kmersLoaded=false;
	private final String outrefstats;
	private final boolean printNonZeroOnly;
	private final boolean overwrite;
	private final int k;
	private final boolean unsupportedAssorted;

	/** Batches smaller than this are aggregated by BBDukGpuBatch before they cross PCIe. */
	static final int MIN_BATCH=1<<20;

	/*--------------------------------------------------------------*/
	/*----------------        Native Methods        ----------------*/
	/*--------------------------------------------------------------*/

	private static native long createNative(int[] cfg);
	private static native int describeNative(int[] cfg, long[] v16);
	private static native int addRefNative(long h, byte[] bases, long[] offsets, int nSeqs);
	/** out = {storedKmers, refKmers} */
	private static native int finalizeNative(long h, long[] out);
	/** bbduk_b200_replicate: out[i] = handle on deviceIds[i] */
	private static native int replicateNative(long h, int[] deviceIds, long[] out);
	private static native long dumpNative(long h, long[] keys, int[] ids);
	private static native int processNative(long h, byte[] bases, long[] offsets, long nReads, boolean paired,
			int[] id0, int[] id0b, int[] lo, int[] hi, byte[] flags, int[] count, int[] maskBits, long[] maskOff, long[] stats8);
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
