package bbduk;

import java.util.ArrayList;

import shared.Tools;
import shared.TrimRead;
import stream.Read;

/**
 * Per-ProcessThread staging between BBDukProcessorS.processList and BBDukIndexGPU (gpu=t).
 *
 * processList calls run(reads) once per list BEFORE its per-pair loop: every read of the list is flattened into
 * one byte[] (mates adjacent), the k-mer block of the whole list is answered by one native call
 * (bbduk_b200_process, include/bbduk_b200.h), and the per-pair loop then calls apply(i, r1, r2, ...) at the
 * place where ktrim() / ktrimTips() / kmask() / ksplit() / countSetKmers() ... used to be called
 * (bbduk/BBDukProcessorS.java:947-1093). Answering the list up front is exact because gpu=t excludes the steps
 * that edit bases ahead of the k-mer block (ftl / ftr / ftr2 / ftm, ecc; BBDukParser checks it) -- reads that the
 * earlier filters discard simply never ask for their answer.
 *
 * Not compiled here (no JDK in the build image); written against the reference's classes as they stand.
 */
final class BBDukGpuBatch {

	BBDukGpuBatch(BBDukParser p, BBDukIndexGPU index_, int tnum){
		index=index_;
		handle=index.handleFor(tnum);
		ktrimN=p.ktrimN;
		ksplit=p.ksplit;
		kfilter=!(p.ktrimLeft || p.ktrimRight || p.ktrimN || p.ksplit);
		kmaskLowercase=p.kmaskLowercase;
		trimSymbol=p.trimSymbol;
	}

	/*--------------------------------------------------------------*/
	/*----------------          One List            ----------------*/
	/*--------------------------------------------------------------*/

	/** Flattens the list, answers its k-mer block on the GPU. Reads are validated exactly as processList would. */
	void run(ArrayList<Read> reads){
		final int n=reads.size();
		paired=(n>0 && reads.get(0).mate!=null);
		final int per=(paired ? 2 : 1);
		long total=0;
		for(Read r1 : reads){
			assert((r1.mate!=null)==paired) : "a list mixes paired and unpaired reads";
			if(!r1.validated()){r1.validate(true);}
			total+=r1.length();
			if(r1.mate!=null){
				if(!r1.mate.validated()){r1.mate.validate(true);}
				total+=r1.mate.length();
			}
		}
		nReads=n*per;
		ensure(nReads, total);
		int pos=0, q=0;
		long words=0;
		for(Read r1 : reads){
			for(int m=0; m<per; m++){
				final Read r=(m==0 ? r1 : r1.mate);
				final int len=r.length();
				if(len>0){System.arraycopy(r.bases, 0, bases, pos, len);}
				pos+=len;
				offsets[++q]=pos;
				if(ktrimN){words+=(len+31)>>>5; maskOff[q]=words;}
			}
		}
		if(ktrimN && (maskBits==null || maskBits.length<words)){maskBits=new int[(int)Tools.max(words, 1024)];}
		java.util.Arrays.fill(stats8, 0);
		if(nReads>0 && !index.processBatch(handle, bases, offsets, nReads, paired, id0, id0b, lo, hi, flags, count,
				ktrimN ? maskBits : null, ktrimN ? maskOff : null, stats8)){
			throw new RuntimeException("bbduk_b200_process failed: "+index.lastError(handle));
		}
	}

	/**
	 * Applies the answer for pair i of the list run() was given; replaces the body of `if(doKmerTrimming){...}else
	 * if(doKmerFiltering){...}` (bbduk/BBDukProcessorS.java:948-1092). Leaves this pair's contribution to the four
	 * k-mer counters in xsum / rktsum / kfReads / kfBases, computed as the reference computes them (:1016-1038,
	 * :1078-1087) -- per pair, because pairs removed by the earlier filters never get here although the device saw them.
	 * @return true if the pair is to be removed (the caller adds it to `bad` under the reference's own conditions)
	 */
	boolean apply(int i, Read r1, Read r2){
		final int a=(paired ? 2*i : i);
		final int len1=r1.length(), len2=(r2==null ? 0 : r2.length());
		xsum=rktsum=kfReads=kfBases=0;
		applyRead(r1, a);
		if(r2!=null){applyRead(r2, a+1);}
		final boolean remove=(ksplit ? r1.mate!=null : (flags[a]&F_REMOVED)!=0);
		if(kfilter){
			if(remove){
				kfReads=(r2==null ? 1 : 2);
				kfBases=len1+len2;
			}
			return remove;
		}
		final int x1=count[a], x2=(r2==null ? 0 : count[a+1]);
		if(ksplit){
			final int trimmed=len1-r1.pairLength();
			xsum=trimmed;
			rktsum=(trimmed>0 ? 1 : 0);
			return remove;
		}
		xsum=x1+x2;
		rktsum=(x1>0 ? 1 : 0)+(x2>0 ? 1 : 0);
		if(remove){
			if(!ktrimN){
				xsum+=(len1-x1)+(len2-x2); //rlen1+rlen2: the lengths right after the scans
				rktsum=(r2==null ? 1 : 2);
			}
		}else if(r2!=null && ((flags[a]|flags[a+1])&F_TPE)!=0){
			//trimpairsevenly cut the longer mate: x = its length after the scan minus what it keeps now
			final int b=((flags[a]&F_TPE)!=0 ? a : a+1);
			final int lenb=(b==a ? len1 : len2);
			final int x=(lenb-count[b])-(hi[b]-lo[b]);
			if(rktsum<2){rktsum++;}
			xsum+=x;
		}
		return remove;
	}

	private void applyRead(Read r, int a){
		final int len=r.length();
		final int f=flags[a];
		if(ktrimN){
			//Replace kmer hit zone with the trim symbol (bbduk/BBDukProcessorS.java:2309-2319)
			final byte[] b=r.bases, quals=r.quality;
			final int w0=(int)maskOff[a];
			for(int j=0; j<len; j++){
				if(((maskBits[w0+(j>>>5)]>>>(j&31))&1)!=0){
					if(kmaskLowercase){
						b[j]=(byte)Tools.toLowerCase(b[j]);
					}else{
						b[j]=trimSymbol;
						if(quals!=null && trimSymbol=='N'){quals[j]=0;}
					}
				}
			}
		}else if(ksplit && (f&F_SPLIT)!=0){
			//bbduk/BBDukProcessorS.java:2482-2489: count[] is rightmost+1, hi[] is leftmost
			final Read r2=r.subRead(count[a], len-1);
			TrimRead.trimByAmount(r, lo[a], len-hi[a], 1, false);
			r.mate=r2;
			r2.mate=r;
			r2.setPairnum(1);
		}else if(!kfilter){
			//ktrim / ktrimTips / ksplit at one end (+ trimpairsevenly): the read keeps original bases [lo, hi)
			final int left=lo[a], right=len-hi[a];
			if(left>0 || right>0){TrimRead.trimByAmount(r, left, right, 1, false);}
		}
		//with trimfailuresto1bp the device already cut the read to one base and cleared the flag, as setDiscarded() does
		if((f&F_DISCARDED)!=0){r.setDiscarded(true);}
	}

	/** id of the scaffold credited for read a of the last list (right tip for ktrim=rl), -1 if none. */
	int id0(int a){return id0[a];}

	private void ensure(int reads, long total){
		assert(total<shared.Shared.MAX_ARRAY_LEN) : "list too large: "+total;
		if(bases==null || bases.length<total){bases=new byte[(int)Tools.min(shared.Shared.MAX_ARRAY_LEN, total+(total>>2)+4096)];}
		if(offsets==null || offsets.length<reads+1){
			final int c=reads+(reads>>2)+1024;
			offsets=new long[c+1];
			maskOff=new long[c+1];
			id0=new int[c];
			id0b=new int[c];
			lo=new int[c];
			hi=new int[c];
			count=new int[c];
			flags=new byte[c];
		}
		offsets[0]=0;
		maskOff[0]=0;
	}

	/*--------------------------------------------------------------*/
	/*----------------            Fields            ----------------*/
	/*--------------------------------------------------------------*/

	private final BBDukIndexGPU index;
	private final long handle;
	private final boolean ktrimN, ksplit, kfilter, kmaskLowercase;
	private final byte trimSymbol;

	private boolean paired;
	private int nReads;
	private byte[] bases;
	private long[] offsets, maskOff;
	private int[] id0, id0b, lo, hi, count, maskBits;
	private byte[] flags;
	private final long[] stats8=new long[8];
	/** Contribution of the pair apply() just handled to basesKTrimmed / readsKTrimmed / readsKFiltered / basesKFiltered */
	int xsum, rktsum, kfReads, kfBases;

	/** flag bits of include/bbduk_b200.h */
	private static final int F_DISCARDED=0x01, F_REMOVED=0x02, F_TPE=0x08, F_SPLIT=0x10;
}
