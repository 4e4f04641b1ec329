// oracle/cs_probe.cpp -- TEST INFRASTRUCTURE ONLY.
//
// Runs the reference's OWN candidate search on a read file and prints every read's candidate list.  Linked against
// the unmodified NextGenMap objects that oracle/Makefile.ngm compiles from /root/reference (everything except
// NGM_main.o): CS::PrefixIteration, CS::PrefixSearch, CS::AddLocationStd, CS::CollectResultsStd, CompactPrefixTable,
// SequenceProvider, ReadProvider are the reference's code; this file only supplies the start-up sequence of
// NGM_main.cpp:123-138 and the per-read reset of CS::RunBatch (CS.cpp:346-392), which cannot be called itself because
// it hands the reads on to the score buffer.
//
// With --bs-mapping / --slam-seq the callback is CS::PrefixMutateSearch, as in CS::RunBatch.
// usage: ngm_cs_probe <ngm options: -r ref -q reads -s sens ...>      (candidates go to stdout, NGM's log to stderr)
// line format: <ReadId> <name> <length> <maxHit> <n> {<location>:<strand>:<votes>}
#include <stdio.h>
#include <cmath>
#include <unistd.h>
#include <sys/stat.h>

#include "NGM.h"
#include "CS.h"

#undef module_name
#define module_name "PROBE"

ILog const * _log = 0;
IConfig * _config = 0;
#ifdef NDEBUG
bool cDebug = false;
#else
bool cDebug = true;
#endif
extern int kCount;

// the one helper the other objects expect from NGM_main.cpp (:344-366): size of a file in bytes
uloc const FileSize(char const * const filename) {
	struct stat st;
	return stat(filename, &st) == 0 ? (uloc) st.st_size : 0;
}

void Help() {
	fprintf(stderr, "ngm_cs_probe: bad arguments\n");
	exit(1);
}

class Probe: public CS {
public:
	Probe() : CS(false) {
	}
	void Run(FILE * out) {
		int const len = 1 << 20;               // the largest table CS::RunBatch ever switches to (CS.cpp:406-407)
		rTable = new CSTableEntry[len];        // CS::DoRun, CS.cpp:466-474
		rList = new int[len];
		for (int i = 0; i < len; ++i) {
			rTable[i].m_Location = (uint) 9223372036854775808u;
			rTable[i].state = -1;
			rList[i] = -1;
		}
		m_CsSensitivity = Config.GetFloat("sensitivity", 0, 1);
		m_RefProvider = NGM.GetRefProvider(0);
		AllocRefEntryChain();
		fprintf(out, "#max_kfreq %d sensitivity %.9g kmer %u\n", maxPrefixFreq = Config.GetInt(MAX_KFREQ), m_CsSensitivity, CS::prefixBasecount);
		while ((m_CurrentBatch = NGM.GetNextReadBatch(m_BatchSize)), (m_CurrentBatch.size() > 0)) {
			for (size_t i = 0; i < m_CurrentBatch.size(); ++i) {
				MappedRead * read = m_CurrentBatch[i];
				m_CurrentSeq = read->ReadId;
				// CS::RunBatch, CS.cpp:341-431, without SendToBuffer: the k-mer callback and the mutated base by mode and mate, the
				// per-read reset, the search in the configured table and its retries in larger ones after an overflow
				PrefixIterationFn pFunc = (m_EnableBS || m_EnableSlamSeq) ? &CS::PrefixMutateSearch : &CS::PrefixSearch;
				ulong mutateFrom, mutateTo;
				if (Config.GetInt("paired") > 0 && (read->ReadId & 1)) {
					mutateFrom = m_EnableSlamSeq ? 0x3 : 0x0;
					mutateTo = m_EnableSlamSeq ? 0x0 : 0x3;
				} else {
					mutateFrom = m_EnableSlamSeq ? 0x1 : 0x2;
					mutateTo = m_EnableSlamSeq ? 0x2 : 0x1;
				}
				m_CurrentReadLength = read->length;
				int const bitsConfigured = c_SrchTableBitLen;
				int bits = bitsConfigured;
				int x = 2;
				bool done = false;
				int tries = 0;
				while (!done && bits <= 20) {
					SetSearchTableBitLen(bits);
					currentState++;
					rListLength = 0;
					maxHitNumber = 0.0f;
					currentThresh = 0.0f;
					if (tries == 0) kCount = 0;
					hpoc = c_SrchTableLen * (tries == 0 ? 0.333f : 0.777f);
					try {
						PrefixIteration(read->Seq, read->length, pFunc, mutateFrom, mutateTo, this, m_PrefixBaseSkip);
						CollectResultsStd(read);
						done = true;
					} catch (int overflow) {
						bits = bitsConfigured + x;
						x += 1;
					}
					++tries;
				}
				SetSearchTableBitLen(bitsConfigured);
				fprintf(out, "%d %s %d %.9g %d", read->ReadId, read->name, read->length, maxHitNumber, read->numScores());
				for (int j = 0; j < read->numScores(); ++j) {
					fprintf(out, " %llu:%d:%.9g", (unsigned long long) read->Scores[j].Location.m_Location, read->Scores[j].Location.isReverse() ? 1 : 0,
							read->Scores[j].Score.f);
				}
				fprintf(out, "\n");
			}
		}
	}
};

int main(int argc, char * argv[]) {
	_NGM::AppName = argv[0];
	InitPlatform();
	_config = new _Config(argc, argv);
	_log = &Log;
	_Log::Init(0, 0);
	NGM;
	NGM.InitProviders();
	Probe * p = new Probe();
	p->Run(stdout);
	fflush(stdout);
	_exit(0);
}
