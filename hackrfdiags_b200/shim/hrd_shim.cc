//**************************************************************************
// file name: hrd_shim.cc
//**************************************************************************
// The reference's eight modulator / demodulator classes re-created on top of
// the libhrd_b200 C ABI (include/hrd.h) as batches of ONE stream, so that
// IqDataProcessor.cc, BasebandDataProcessor.cc, Radio.cc and the stand-alone
// test programs (am.cc, fm.cc, wbfm.cc, ssb.cc) compile and link unchanged.
//
// What each member replaces (reference paths relative to radioDiags/):
//   <X>Demodulator::acceptIqData      AmDemodulator.cc:297-315, FmDemodulator.cc:353-371,
//                                     WbFmDemodulator.cc:341-356, SsbDemodulator.cc:420-438
//   <X>Demodulator::resetDemodulator  AmDemodulator.cc:232-265, FmDemodulator.cc:290-321,
//                                     WbFmDemodulator.cc:265-297, SsbDemodulator.cc:297-331
//   <X>Demodulator::setDemodulatorGain  AmDemodulator.cc:267, FmDemodulator.cc:323, WbFmDemodulator.cc:299, SsbDemodulator.cc:390
//   <X>Modulator::acceptData          AmModulator.cc:366-381, FmModulator.cc:373-388,
//                                     WbFmModulator.cc:347-365, SsbModulator.cc:455-470
//   setModulationIndex / setFrequencyDeviation / set{Lsb,Usb}ModulationMode and their guards
//                                     AmModulator.cc:329-339, FmModulator.cc:336-346,
//                                     WbFmModulator.cc:310-328, SsbModulator.cc:392-446
// The objects keep the reference's threading rule: one data thread per object;
// setters may arrive from another thread and take effect at the next call.
//
// No signal processing happens here.  Without a usable B200 the constructors
// print the library's error and abort: there is deliberately no CPU fallback.
//**************************************************************************
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../../include/hrd.h"

#include "AmDemodulator.h"
#include "FmDemodulator.h"
#include "WbFmDemodulator.h"
#include "SsbDemodulator.h"
#include "AmModulator.h"
#include "FmModulator.h"
#include "WbFmModulator.h"
#include "SsbModulator.h"
#include "IqDataProcessor.h"

// supplied by the host application, exactly as for the reference classes
// (diagUi.cc:2881; the test programs define their own, am.cc:93-110)
extern void nprintf(FILE *s,const char *formatPtr, ...);

static int shimDevice(void)
{
  const char *e = getenv("HRD_DEVICE");
  return (e != NULL) ? atoi(e) : 0;
}

static void shimDie(const char *what)
{
  fprintf(stderr,"libhrdshim: %s failed: %s\n",what,hrd_last_error());
  fprintf(stderr,"libhrdshim: this build has no CPU fallback (needs an NVIDIA B200, sm_100)\n");
  abort();
}

//**************************************************************************
// Receive side.
//**************************************************************************
struct HrdShimRx
{
  hrd_batch_t *batchPtr;
  int mode;
  int gainParameter;
  int resetUnit;
  float demodulatorGain;
  bool lsbDemodulationMode;
  unsigned resetCount; // resetDemodulator() calls so far (an attached IqDataProcessor follows them)
  void (*pcmCallbackPtr)(int16_t *bufferPtr,uint32_t bufferLength);

  // acceptIqData() hands whole PCM samples (64 bytes of IQ at 256000 S/s) to
  // the library; a shorter tail waits here for the next call.  The reference's
  // own caller always passes 32768 bytes, so this never triggers there.
  std::vector<int8_t> pending;
  std::vector<int8_t> work;
  std::vector<int16_t> pcmData;
};

static HrdShimRx *rxCreate(int mode,int gainParameter,int resetUnit,
    void (*pcmCallbackPtr)(int16_t *bufferPtr,uint32_t bufferLength))
{
  HrdShimRx *p = new HrdShimRx;
  p->batchPtr = NULL;
  p->mode = mode;
  p->gainParameter = gainParameter;
  p->resetUnit = resetUnit;
  p->lsbDemodulationMode = true;
  p->resetCount = 0;
  p->pcmCallbackPtr = pcmCallbackPtr;
  if (hrd_create(shimDevice(),1,HRD_RX,&p->batchPtr) != HRD_OK) shimDie("hrd_create");
  if (hrd_set_mode(p->batchPtr,0,mode) != HRD_OK) shimDie("hrd_set_mode");
  if (hrd_get_param(p->batchPtr,0,gainParameter,&p->demodulatorGain) != HRD_OK)
    shimDie("hrd_get_param");
  return (p);
}

static void rxDestroy(HrdShimRx *p)
{
  if (p != NULL)
  {
    hrd_destroy(p->batchPtr);
    delete p;
  }
}

static void rxAccept(HrdShimRx *p,int8_t *bufferPtr,uint32_t bufferLength)
{
  const int8_t *src = bufferPtr;
  size_t total = bufferLength;

  if (!p->pending.empty())
  {
    p->work.assign(p->pending.begin(),p->pending.end());
    p->work.insert(p->work.end(),bufferPtr,bufferPtr + bufferLength);
    src = p->work.data();
    total = p->work.size();
  }

  const size_t whole = total - (total % 64);
  const size_t sampleCount = whole / 64;

  if (p->pcmData.size() < sampleCount + 1) p->pcmData.resize(sampleCount + 1);

  if (whole > 0)
  {
    if (hrd_rx_process(p->batchPtr,src,whole,whole,HRD_ENTRY_256K,p->pcmData.data(),
                       p->pcmData.size(),NULL,HRD_MEM_HOST,NULL) != HRD_OK)
      shimDie("hrd_rx_process");
  }

  p->pending.assign(src + whole,src + total);

  // The reference calls back once per acceptIqData(), even with 0 samples.
  if (p->pcmCallbackPtr != NULL)
    p->pcmCallbackPtr(p->pcmData.data(),(uint32_t)sampleCount);
}

static void rxSetGain(HrdShimRx *p,float gain)
{
  p->demodulatorGain = gain;
  if (hrd_set_param(p->batchPtr,0,p->gainParameter,gain) != HRD_OK) shimDie("hrd_set_param");
}

static void rxReset(HrdShimRx *p)
{
  p->pending.clear();
  p->resetCount++;
  if (hrd_reset(p->batchPtr,0,p->resetUnit) != HRD_OK) shimDie("hrd_reset");
}

#define HRD_SHIM_DEMODULATOR(Class,Title,Mode,Gain,Unit)                          \
  Class::Class(void (*pcmCallbackPtr)(int16_t *bufferPtr,uint32_t bufferLength))  \
  { implPtr = rxCreate(Mode,Gain,Unit,pcmCallbackPtr); }                          \
  Class::~Class(void) { rxDestroy(implPtr); }                                     \
  void Class::resetDemodulator(void) { rxReset(implPtr); }                        \
  void Class::setDemodulatorGain(float gain) { rxSetGain(implPtr,gain); }         \
  void Class::acceptIqData(int8_t *bufferPtr,uint32_t bufferLength)               \
  { rxAccept(implPtr,bufferPtr,bufferLength); }

HRD_SHIM_DEMODULATOR(AmDemodulator,"AM",HRD_MODE_AM,HRD_PARAM_AM_GAIN,HRD_UNIT_AM)
HRD_SHIM_DEMODULATOR(FmDemodulator,"FM",HRD_MODE_FM,HRD_PARAM_FM_GAIN,HRD_UNIT_FM)
HRD_SHIM_DEMODULATOR(WbFmDemodulator,"Wideband FM",HRD_MODE_WBFM,HRD_PARAM_WBFM_GAIN,HRD_UNIT_WBFM)
HRD_SHIM_DEMODULATOR(SsbDemodulator,"SSB",HRD_MODE_LSB,HRD_PARAM_SSB_GAIN,HRD_UNIT_SSB)

void SsbDemodulator::setLsbDemodulationMode(void)
{
  implPtr->lsbDemodulationMode = true;
  if (hrd_set_mode(implPtr->batchPtr,0,HRD_MODE_LSB) != HRD_OK) shimDie("hrd_set_mode");
}

void SsbDemodulator::setUsbDemodulationMode(void)
{
  implPtr->lsbDemodulationMode = false;
  if (hrd_set_mode(implPtr->batchPtr,0,HRD_MODE_USB) != HRD_OK) shimDie("hrd_set_mode");
}

// Same text as the reference prints (AmDemodulator.cc:553-563 and siblings).
static void rxDisplay(const char *titlePtr,float gain)
{
  nprintf(stderr,"\n--------------------------------------------\n");
  nprintf(stderr,"%s Demodulator Internal Information\n",titlePtr);
  nprintf(stderr,"--------------------------------------------\n");
  nprintf(stderr,"Demodulator Gain         : %f\n",gain);
}

void AmDemodulator::displayInternalInformation(void) { rxDisplay("AM",implPtr->demodulatorGain); }
void FmDemodulator::displayInternalInformation(void) { rxDisplay("FM",implPtr->demodulatorGain); }
void WbFmDemodulator::displayInternalInformation(void) { rxDisplay("Wideband FM",implPtr->demodulatorGain); }

void SsbDemodulator::displayInternalInformation(void)
{
  nprintf(stderr,"\n--------------------------------------------\n");
  nprintf(stderr,"SSB Demodulator Internal Information\n");
  nprintf(stderr,"--------------------------------------------\n");
  nprintf(stderr,"Demodulation Mode        : ");
  if (implPtr->lsbDemodulationMode)
  {
    nprintf(stderr,"LSB\n");
  }
  else
  {
    nprintf(stderr,"USB\n");
  }
  nprintf(stderr,"Demodulator Gain         : %f\n",implPtr->demodulatorGain);
}

//**************************************************************************
// Transmit side.
//**************************************************************************
struct HrdShimTx
{
  hrd_batch_t *batchPtr;
  int mode;
  int resetUnit;
  bool lsbModulationMode;
};

static HrdShimTx *txCreate(int mode,int resetUnit)
{
  HrdShimTx *p = new HrdShimTx;
  p->batchPtr = NULL;
  p->mode = mode;
  p->resetUnit = resetUnit;
  p->lsbModulationMode = true;
  if (hrd_create(shimDevice(),1,HRD_TX,&p->batchPtr) != HRD_OK) shimDie("hrd_create");
  if (hrd_set_mode(p->batchPtr,0,mode) != HRD_OK) shimDie("hrd_set_mode");
  return (p);
}

static void txDestroy(HrdShimTx *p)
{
  if (p != NULL)
  {
    hrd_destroy(p->batchPtr);
    delete p;
  }
}

static void txAccept(HrdShimTx *p,int16_t *bufferPtr,uint32_t bufferLength,
                     int8_t *outputBufferPtr,uint32_t *outputBufferLengthPtr)
{
  // bufferLength PCM samples -> bufferLength * 256 IQ samples = * 512 bytes,
  // whatever the caller's buffer size argument said (AmModulator.cc:410-530)
  const size_t outputLength = (size_t)bufferLength * 512;
  if (bufferLength > 0)
  {
    if (hrd_tx_process(p->batchPtr,bufferPtr,bufferLength,bufferLength,outputBufferPtr,
                       outputLength,HRD_MEM_HOST,NULL) != HRD_OK)
      shimDie("hrd_tx_process");
  }
  *outputBufferLengthPtr = (uint32_t)outputLength;
}

static float txParameter(HrdShimTx *p,int parameter)
{
  float value = 0;
  if (hrd_get_param(p->batchPtr,0,parameter,&value) != HRD_OK) shimDie("hrd_get_param");
  return (value);
}

#define HRD_SHIM_MODULATOR(Class,Mode,Unit)                                       \
  Class::Class(void) { implPtr = txCreate(Mode,Unit); }                           \
  Class::~Class(void) { txDestroy(implPtr); }                                     \
  void Class::resetModulator(void)                                                \
  { if (hrd_reset(implPtr->batchPtr,0,Unit) != HRD_OK) shimDie("hrd_reset"); }    \
  void Class::acceptData(int16_t *bufferPtr,uint32_t bufferLength,                \
                         int8_t *outputBufferPtr,uint32_t *outputBufferLengthPtr) \
  { txAccept(implPtr,bufferPtr,bufferLength,outputBufferPtr,outputBufferLengthPtr); }

HRD_SHIM_MODULATOR(AmModulator,HRD_MODE_AM,HRD_UNIT_AM)
HRD_SHIM_MODULATOR(FmModulator,HRD_MODE_FM,HRD_UNIT_FM)
HRD_SHIM_MODULATOR(WbFmModulator,HRD_MODE_WBFM,HRD_UNIT_WBFM)
HRD_SHIM_MODULATOR(SsbModulator,HRD_MODE_LSB,HRD_UNIT_SSB)

// The range checks (and the reference's quirk of testing the OLD deviation,
// FmModulator.cc:336-346) live behind hrd_set_param so that batched callers
// get them too.
void AmModulator::setModulationIndex(float modulationIndex)
{
  if (hrd_set_param(implPtr->batchPtr,0,HRD_PARAM_AM_INDEX,modulationIndex) != HRD_OK)
    shimDie("hrd_set_param");
}

void FmModulator::setFrequencyDeviation(float deviation)
{
  if (hrd_set_param(implPtr->batchPtr,0,HRD_PARAM_FM_DEV,deviation) != HRD_OK)
    shimDie("hrd_set_param");
}

void WbFmModulator::setFrequencyDeviation(float deviation)
{
  if (hrd_set_param(implPtr->batchPtr,0,HRD_PARAM_WBFM_DEV,deviation) != HRD_OK)
    shimDie("hrd_set_param");
}

void SsbModulator::setLsbModulationMode(void)
{
  implPtr->lsbModulationMode = true;
  if (hrd_set_mode(implPtr->batchPtr,0,HRD_MODE_LSB) != HRD_OK) shimDie("hrd_set_mode");
}

void SsbModulator::setUsbModulationMode(void)
{
  implPtr->lsbModulationMode = false;
  if (hrd_set_mode(implPtr->batchPtr,0,HRD_MODE_USB) != HRD_OK) shimDie("hrd_set_mode");
}

void AmModulator::displayInternalInformation(void)
{
  nprintf(stderr,"\n--------------------------------------------\n");
  nprintf(stderr,"AM Modulator Internal Information\n");
  nprintf(stderr,"--------------------------------------------\n");
  nprintf(stderr,"Modulator Index          : %f\n",txParameter(implPtr,HRD_PARAM_AM_INDEX));
}

void FmModulator::displayInternalInformation(void)
{
  nprintf(stderr,"\n--------------------------------------------\n");
  nprintf(stderr,"FM Modulator Internal Information\n");
  nprintf(stderr,"--------------------------------------------\n");
  nprintf(stderr,"Frequency Deviation:      : %fHz\n",txParameter(implPtr,HRD_PARAM_FM_DEV));
}

void WbFmModulator::displayInternalInformation(void)
{
  nprintf(stderr,"\n--------------------------------------------\n");
  nprintf(stderr,"Wideband FM Modulator Internal Information\n");
  nprintf(stderr,"--------------------------------------------\n");
  nprintf(stderr,"Frequency Deviation:      : %fHz\n",txParameter(implPtr,HRD_PARAM_WBFM_DEV));
}

void SsbModulator::displayInternalInformation(void)
{
  nprintf(stderr,"\n--------------------------------------------\n");
  nprintf(stderr,"SSB Modulator Internal Information\n");
  nprintf(stderr,"--------------------------------------------\n");
  nprintf(stderr,"Modulation Mode        : ");
  if (implPtr->lsbModulationMode)
  {
    nprintf(stderr,"LSB\n");
  }
  else
  {
    nprintf(stderr,"USB\n");
  }
}

//**************************************************************************
// IqDataProcessor (radioDiags/src_diags/IqDataProcessor.cc): the 2.048 MS/s entry.
//**************************************************************************
// One stream of an Rx batch at HRD_ENTRY_2048K does what acceptIqData does
// (IqDataProcessor.cc:926-1038): three half-band decimators, the Fs/4 rotation,
// Squelch::run on the 256 kS/s block, the gated demodulator.  The demodulator
// OBJECTS the application hands in stay the owners of their parameters and PCM
// callbacks: before every block the active one's gain and reset count are
// mirrored into this object's batch, and its callback receives the PCM -- once
// per block the squelch lets through, exactly when the reference would have
// called <X>Demodulator::acceptIqData.
// The UDP dump of the 256 kS/s stream (IqDataProcessor.cc:953-957) is network
// I/O and stays out: the enable flag is kept and reported, nothing is sent.

// the receive gain Squelch::run refers the level to (Radio.cc:15,413,1599)
extern uint32_t radio_adjustableReceiveGainInDb;

struct HrdShimIqdp
{
  hrd_batch_t *batchPtr;      // acceptIqData: the whole chain
  hrd_batch_t *frontEndPtr;   // reduceSampleRate() called on its own
  IqDataProcessor::demodulatorType demodulatorMode;
  int32_t signalDetectThreshold;
  HrdShimRx *demodulator[4];  // AM, FM, WBFM, SSB as handed in
  float pushedGain[4];
  unsigned seenReset[4];
  uint32_t pushedReceiveGain;
  unsigned long blockBytes;
  bool iqDumpEnabled;
  bool signalNotificationEnabled;
  void *signalCallbackContextPtr;
  void (*signalCallbackPtr)(bool signalPresent,void *contextPtr);
  bool signalMagnitudeNotificationEnabled;
  void *signalMagnitudeCallbackContextPtr;
  void (*signalMagnitudeCallbackPtr)(uint32_t signalMagnitude,void *contextPtr);
  std::vector<int16_t> pcmData;
};

IqDataProcessor::IqDataProcessor(char *hostIpAddress,int hostPort)
{
  (void)hostIpAddress;
  (void)hostPort;
  HrdShimIqdp *p = new HrdShimIqdp;
  p->batchPtr = NULL;
  p->frontEndPtr = NULL;
  p->demodulatorMode = None;        // IqDataProcessor.cc:70
  p->signalDetectThreshold = -200;  // IqDataProcessor.cc:120
  for (int i = 0; i < 4; i++)
  {
    p->demodulator[i] = NULL;
    p->pushedGain[i] = 0;
    p->seenReset[i] = 0;
  }
  p->pushedReceiveGain = 16;
  p->blockBytes = 0;
  p->iqDumpEnabled = false;
  p->signalNotificationEnabled = false;
  p->signalCallbackPtr = NULL;
  p->signalCallbackContextPtr = NULL;
  p->signalMagnitudeNotificationEnabled = false;
  p->signalMagnitudeCallbackPtr = NULL;
  p->signalMagnitudeCallbackContextPtr = NULL;
  if (hrd_create(shimDevice(),1,HRD_RX,&p->batchPtr) != HRD_OK) shimDie("hrd_create");
  // always the squelched path: the magnitude and the decision of every block are part of the interface
  if (hrd_set_option(p->batchPtr,HRD_OPT_RX_SQUELCH,1) != HRD_OK) shimDie("hrd_set_option");
  memset(decimatedData,0,sizeof decimatedData);
  implPtr = p;
}

IqDataProcessor::~IqDataProcessor(void)
{
  hrd_destroy(implPtr->batchPtr);
  hrd_destroy(implPtr->frontEndPtr);
  delete implPtr;
}

void IqDataProcessor::setAmDemodulator(AmDemodulator *demodulatorPtr) { implPtr->demodulator[0] = demodulatorPtr->implPtr; }
void IqDataProcessor::setFmDemodulator(FmDemodulator *demodulatorPtr) { implPtr->demodulator[1] = demodulatorPtr->implPtr; }
void IqDataProcessor::setWbFmDemodulator(WbFmDemodulator *demodulatorPtr) { implPtr->demodulator[2] = demodulatorPtr->implPtr; }
void IqDataProcessor::setSsbDemodulator(SsbDemodulator *demodulatorPtr) { implPtr->demodulator[3] = demodulatorPtr->implPtr; }

// IqDataProcessor.cc:346-372: LSB / USB also flip the SSB demodulator object's sideband
void IqDataProcessor::setDemodulatorMode(demodulatorType mode)
{
  implPtr->demodulatorMode = mode;
  if (implPtr->demodulator[3] != NULL)
  {
    if (mode == Lsb) implPtr->demodulator[3]->lsbDemodulationMode = true;
    if (mode == Usb) implPtr->demodulator[3]->lsbDemodulationMode = false;
  }
  if (hrd_set_mode(implPtr->batchPtr,0,(int)mode) != HRD_OK) shimDie("hrd_set_mode");
}

// IqDataProcessor.cc:392-405
void IqDataProcessor::setSignalDetectThreshold(int32_t threshold)
{
  implPtr->signalDetectThreshold = threshold;
  if (hrd_set_param(implPtr->batchPtr,0,HRD_PARAM_SQUELCH_THRESHOLD,(float)threshold) != HRD_OK)
    shimDie("hrd_set_param");
}

void IqDataProcessor::enableSignalNotification(void) { implPtr->signalNotificationEnabled = true; }
void IqDataProcessor::disableSignalNotification(void) { implPtr->signalNotificationEnabled = false; }
void IqDataProcessor::registerSignalStateCallback(void (*callbackPtr)(bool signalPresent,void *contextPtr),void *contextPtr)
{
  implPtr->signalCallbackContextPtr = contextPtr;
  implPtr->signalCallbackPtr = callbackPtr;
}
void IqDataProcessor::enableSignalMagnitudeNotification(void) { implPtr->signalMagnitudeNotificationEnabled = true; }
void IqDataProcessor::disableSignalMagnitudeNotification(void) { implPtr->signalMagnitudeNotificationEnabled = false; }
void IqDataProcessor::registerSignalMagnitudeCallback(void (*callbackPtr)(uint32_t signalMagnitude,void *contextPtr),void *contextPtr)
{
  implPtr->signalMagnitudeCallbackContextPtr = contextPtr;
  implPtr->signalMagnitudeCallbackPtr = callbackPtr;
}
void IqDataProcessor::enableIqDump(void) { implPtr->iqDumpEnabled = true; }
void IqDataProcessor::disableIqDump(void) { implPtr->iqDumpEnabled = false; }
bool IqDataProcessor::isIqDumpEnabled(void) { return (implPtr->iqDumpEnabled); }

// IqDataProcessor.cc:926-1038
void IqDataProcessor::acceptIqData(unsigned long timeStamp,int8_t *bufferPtr,unsigned long byteCount)
{
  (void)timeStamp;
  HrdShimIqdp *p = implPtr;
  static const int gainParameter[4] = {HRD_PARAM_AM_GAIN,HRD_PARAM_FM_GAIN,HRD_PARAM_WBFM_GAIN,HRD_PARAM_SSB_GAIN};
  static const int resetUnit[4] = {HRD_UNIT_AM,HRD_UNIT_FM,HRD_UNIT_WBFM,HRD_UNIT_SSB};
  static const int slotOfMode[6] = {-1,0,1,2,3,3};

  byteCount -= byteCount % 512; // whole PCM samples (the reference's caller always passes 262144)
  if (byteCount == 0) return;

  // one call is one squelch decision
  if (byteCount != p->blockBytes)
  {
    if (hrd_set_option(p->batchPtr,HRD_OPT_RX_SQUELCH_BLOCK,(int)byteCount) != HRD_OK) shimDie("hrd_set_option");
    p->blockBytes = byteCount;
  }
  if (radio_adjustableReceiveGainInDb != p->pushedReceiveGain)
  {
    if (hrd_set_param(p->batchPtr,0,HRD_PARAM_RX_GAIN_DB,(float)radio_adjustableReceiveGainInDb) != HRD_OK)
      shimDie("hrd_set_param");
    p->pushedReceiveGain = radio_adjustableReceiveGainInDb;
  }
  // the demodulator objects own their parameters: follow them
  for (int i = 0; i < 4; i++)
  {
    HrdShimRx *d = p->demodulator[i];
    if (d == NULL) continue;
    if (d->demodulatorGain != p->pushedGain[i])
    {
      if (hrd_set_param(p->batchPtr,0,gainParameter[i],d->demodulatorGain) != HRD_OK) shimDie("hrd_set_param");
      p->pushedGain[i] = d->demodulatorGain;
    }
    if (d->resetCount != p->seenReset[i])
    {
      if (hrd_reset(p->batchPtr,0,resetUnit[i]) != HRD_OK) shimDie("hrd_reset");
      p->seenReset[i] = d->resetCount;
    }
  }

  const size_t sampleCount = byteCount / 512;
  if (p->pcmData.size() < sampleCount + 1) p->pcmData.resize(sampleCount + 1);
  uint32_t produced = 0;
  if (hrd_rx_process(p->batchPtr,bufferPtr,byteCount,byteCount,HRD_ENTRY_2048K,p->pcmData.data(),
                     p->pcmData.size(),&produced,HRD_MEM_HOST,NULL) != HRD_OK)
    shimDie("hrd_rx_process");
  uint32_t magnitude = 0, blocks = 0;
  uint8_t allowed = 0;
  if (hrd_rx_squelch_report(p->batchPtr,&magnitude,&allowed,1,&blocks) != HRD_OK) shimDie("hrd_rx_squelch_report");

  if (p->signalNotificationEnabled && (p->signalCallbackPtr != NULL))
    p->signalCallbackPtr(allowed != 0,p->signalCallbackContextPtr);
  if (p->signalMagnitudeNotificationEnabled && (p->signalMagnitudeCallbackPtr != NULL))
    p->signalMagnitudeCallbackPtr(magnitude,p->signalMagnitudeCallbackContextPtr);

  if (allowed)
  {
    const int slot = slotOfMode[(int)p->demodulatorMode];
    if ((slot >= 0) && (p->demodulator[slot] != NULL) && (p->demodulator[slot]->pcmCallbackPtr != NULL))
      p->demodulator[slot]->pcmCallbackPtr(p->pcmData.data(),produced);
  }
}

// reduceSampleRate on its own (IqDataProcessor.cc:429-500): the front end of a second one-stream batch,
// with the rotation the library fuses into it taken back out, so that decimatedData holds what the
// reference's holds at this point.  (acceptIqData does not come through here.)
uint32_t IqDataProcessor::reduceSampleRate(int8_t *bufferPtr,uint32_t bufferLength)
{
  HrdShimIqdp *p = implPtr;
  if (p->frontEndPtr == NULL)
  {
    if (hrd_create(shimDevice(),1,HRD_RX,&p->frontEndPtr) != HRD_OK) shimDie("hrd_create");
  }
  bufferLength -= bufferLength % 512;
  if (bufferLength > 262144) bufferLength = 262144;
  if (bufferLength == 0) return (0);
  if (hrd_rx_front_end(p->frontEndPtr,bufferPtr,bufferLength,bufferLength,decimatedData,sizeof decimatedData,
                       HRD_MEM_HOST,NULL) != HRD_OK)
    shimDie("hrd_rx_front_end");
  downconvertByFsOver4(decimatedData,bufferLength / 8);
  return (bufferLength / 8);
}

// IqDataProcessor.cc:771-815 and 715-759: the Fs/4 rotations on their own, in place (hrd_rx_fs4_rotate)
static void shimRotate(HrdShimIqdp *p,int8_t *bufferPtr,uint32_t byteCount,int up)
{
  if (p->frontEndPtr == NULL)
  {
    if (hrd_create(shimDevice(),1,HRD_RX,&p->frontEndPtr) != HRD_OK) shimDie("hrd_create");
  }
  byteCount -= byteCount % 8;
  if (hrd_rx_fs4_rotate(p->frontEndPtr,bufferPtr,byteCount,up,HRD_MEM_HOST,NULL) != HRD_OK) shimDie("hrd_rx_fs4_rotate");
}

void IqDataProcessor::upconvertByFsOver4(int8_t *bufferPtr,uint32_t byteCount) { shimRotate(implPtr,bufferPtr,byteCount,1); }
void IqDataProcessor::downconvertByFsOver4(int8_t *bufferPtr,uint32_t byteCount) { shimRotate(implPtr,bufferPtr,byteCount,0); }

// IqDataProcessor.cc:1058-1124
void IqDataProcessor::displayInternalInformation(void)
{
  static const char *modeName[6] = {"None","AM","FM","WBFM","LSB","USB"};
  nprintf(stderr,"\n--------------------------------------------\n");
  nprintf(stderr,"IQ Data Processor Internal Information\n");
  nprintf(stderr,"--------------------------------------------\n");
  nprintf(stderr,"Demodulator Mode         : ");
  if (((int)implPtr->demodulatorMode >= 0) && ((int)implPtr->demodulatorMode <= 5))
    nprintf(stderr,"%s\n",modeName[(int)implPtr->demodulatorMode]);
  nprintf(stderr,"Signal Detect Threhold   : %d dBFs\n",implPtr->signalDetectThreshold);
  if (implPtr->iqDumpEnabled)
  {
    nprintf(stderr,"IQ Dump Enabled          : Yes\n");
  }
  else
  {
    nprintf(stderr,"IQ Dump Enabled          : No\n");
  }
}
