// dsp.h -- data contracts of the module / flow interface.
//
// Mirror of the reference's operator interface (cudarecv/dsp/inc/dsp.h:20-146): the same
// enumerator names and the same Port / Param / ExpectedPort fields, so that flow code written
// against the reference (dsp/src/dpeflow.cpp) reads the same here.  Differences, on purpose:
// Port carries a 64-bit Length next to the 16-bit VectorLength the reference is limited to
// (10 MHz x 20 ms = 200 000 samples overflows an unsigned short, sampleblock.h:81).
#ifndef DPE_HOST_DSP_H_
#define DPE_HOST_DSP_H_

#include <cstdint>
#include <cstddef>

namespace dsp {

const unsigned short VECTORLENGTH_ANY = 0;

enum ValueType_t : uint8_t {
    VALUETYPE_ANY = 0, VALUE, VALUE_CMPX, RATIO, RATIO_DB, FREQUENCY_HZ, FREQUENCY_RAD, PHASE, MAGNITUDE,
    RS_CORR_OUT, SS_CORR_OUT, CHANNEL, STATE, COVARIANCE, FUNCTION_PTR, EPHEMS, GRID
};

enum MemLoc_t : uint8_t { HOST = 0, CUDA_DEVICE = 1 };

enum DataType_t : uint8_t {
    DATATYPE_ANY = 0, UNDEFINED_t, FLOAT_t, DOUBLE_t, FIXED_Q15_t, FIXED_Q31_t, FIXED_I15Q16_t, CHAR_t,
    STRING_t, INT_t, BOOL_t, CUFFTCOMP_t
};

struct Param {
    void* Ptr;
    DataType_t Datatype;
    unsigned int Capacity;   // bytes
    unsigned int Size;       // bytes
};

struct Port {
    char Name[32];
    DataType_t Datatype;
    signed char Exponent;
    ValueType_t ValueType;
    MemLoc_t MemLoc;
    unsigned short VectorLength;   // as in the reference (saturates at 65535)
    void* Data;
    int AuxValue;
    int64_t Length;                // true element count
};

struct ExpectedPort {
    char Name[32];
    DataType_t Datatype;
    ValueType_t ValueType;
    unsigned short VectorLength;   // 1 = scalar, 0 = any length
};

class Module;
class Flow;
class FlowMgr;
class DPEFlow;
class DPInit;
class SampleBlock;
class BatchCorrScores;
class BatchCorrManifold;
class cuEKF;
class cuChanMgr;
class DataLogger;

}  // namespace dsp
#endif
