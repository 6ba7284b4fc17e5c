// C ABI, layer C: the JNI symbols Kompressor's JVM binding loads from "zstd-jni" (libzstd-jni.so),
// decompress side, re-pointed at the kzg streaming shim (layer B).  Compiled only where <jni.h> exists
// (kompressor_b200/build.py: build_jni); nothing else in this repository depends on it.
//
// Replaces, symbol for symbol (reference: kompressor-zstd--nativelib/src/jvmCommonMain/jni/Wrapper.cpp):
//   Java_com_ensody_kompressor_zstd_ZstdWrapper_createDecompressor          Wrapper.cpp:123-130
//   Java_com_ensody_kompressor_zstd_ZstdWrapper_freeDecompressor            Wrapper.cpp:132-140
//   Java_com_ensody_kompressor_zstd_ZstdWrapper_loadDecompressorDictionary  Wrapper.cpp:58-73
//   Java_com_ensody_kompressor_zstd_ZstdWrapper_decompressStream            Wrapper.cpp:142-187
//   Java_com_ensody_kompressor_zstd_ZstdWrapper_getErrorName                Wrapper.cpp:189-196
//   JNI_OnLoad / JNI_OnUnload (field IDs of ByteArraySlice)                 jni/common/src/DefaultLoad.cpp:11-37,
//                                                                           jni/common/src/SliceClass.cpp:4-7
//   Java_com_ensody_kompressor_zstd_ZstdWrapper_createCompressor / freeCompressor   Wrapper.cpp:10-28
//   Java_com_ensody_kompressor_zstd_ZstdWrapper_setParameter                        Wrapper.cpp:30-40
//   Java_com_ensody_kompressor_zstd_ZstdWrapper_loadCompressorDictionary            Wrapper.cpp:42-56
//   Java_com_ensody_kompressor_zstd_ZstdWrapper_compressStream                      Wrapper.cpp:75-121
// The Kotlin declarations they bind are ZstdWrapper.kt:24-60 (a Kotlin `object`, hence the jobject
// receiver).  With the compress-side symbols (on top of the batch frame writer) the library exports
// everything libzstd-jni.so does; a build that keeps libzstd for compression simply does not re-point
// those five (INTEGRATION.md shows the two libraries loaded side by side).
//
// Difference from the reference on purpose: array access is limited to
// the slice ranges (Get/SetByteArrayRegion) instead of Get/ReleaseByteArrayElements on the whole arrays,
// so a call moves at most the bytes it can consume or produce.
#include <jni.h>
#include <vector>

#include "../../include/kzg.h"

namespace {
jfieldID g_readStart = nullptr, g_writeStart = nullptr;

bool resolve_fields(JNIEnv* env) {
    if (g_readStart && g_writeStart) return true;
    jclass cls = env->FindClass("com/ensody/kompressor/core/ByteArraySlice");
    if (!cls) return false;
    g_readStart = env->GetFieldID(cls, "readStart", "I");
    g_writeStart = env->GetFieldID(cls, "writeStart", "I");
    return g_readStart && g_writeStart;
}
const jlong kGenericError = -1;      // -ZSTD_error_GENERIC, what the reference returns on array-access failure
}  // namespace

extern "C" {

JNIEXPORT jint JNICALL JNI_OnLoad(JavaVM* vm, void*) {
    JNIEnv* env = nullptr;
    if (vm->GetEnv(reinterpret_cast<void**>(&env), JNI_VERSION_1_4) != JNI_OK) return JNI_ERR;
    if (!resolve_fields(env)) return JNI_ERR;
    return env->GetVersion();
}

JNIEXPORT void JNICALL JNI_OnUnload(JavaVM*, void*) { g_readStart = g_writeStart = nullptr; }

JNIEXPORT jlong JNICALL Java_com_ensody_kompressor_zstd_ZstdWrapper_createDecompressor(JNIEnv*, jobject) {
    return reinterpret_cast<jlong>(kzg_createDCtx());         // 0 if no GPU: Kotlin throws "Failed allocating zstd dctx"
}

JNIEXPORT void JNICALL Java_com_ensody_kompressor_zstd_ZstdWrapper_freeDecompressor(JNIEnv*, jobject, jlong dctx) {
    kzg_freeDCtx(reinterpret_cast<kzg_DCtx*>(dctx));           // arrives from the Cleaner daemon thread
}

JNIEXPORT jlong JNICALL Java_com_ensody_kompressor_zstd_ZstdWrapper_loadDecompressorDictionary(JNIEnv* env, jobject, jlong dctx,
                                                                                               jbyteArray dictionary) {
    const jsize n = env->GetArrayLength(dictionary);
    std::vector<jbyte> buf(static_cast<size_t>(n > 0 ? n : 0));
    if (n > 0) {
        env->GetByteArrayRegion(dictionary, 0, n, buf.data());
        if (env->ExceptionCheck()) return kGenericError;
    }
    return static_cast<jlong>(kzg_DCtx_loadDictionary(reinterpret_cast<kzg_DCtx*>(dctx), buf.data(), buf.size()));
}

JNIEXPORT jlong JNICALL Java_com_ensody_kompressor_zstd_ZstdWrapper_decompressStream(
        JNIEnv* env, jobject, jlong dctx, jobject inputSlice, jbyteArray inputByteArray, jint inputStart, jint inputEndExclusive,
        jobject outputSlice, jbyteArray outputByteArray, jint outputStart, jint outputEndExclusive) {
    if (!resolve_fields(env)) return kGenericError;
    const jsize in_n = inputEndExclusive > inputStart ? inputEndExclusive - inputStart : 0;
    const jsize out_n = outputEndExclusive > outputStart ? outputEndExclusive - outputStart : 0;
    std::vector<jbyte> in(static_cast<size_t>(in_n)), out(static_cast<size_t>(out_n));
    if (in_n > 0) {
        env->GetByteArrayRegion(inputByteArray, inputStart, in_n, in.data());
        if (env->ExceptionCheck()) return kGenericError;
    }
    // positions are relative to the slice ranges here; the cursors written back are absolute again
    kzg_inBuffer ib = { in.data(), static_cast<size_t>(in_n), 0 };
    kzg_outBuffer ob = { out.data(), static_cast<size_t>(out_n), 0 };
    const size_t result = kzg_decompressStream(reinterpret_cast<kzg_DCtx*>(dctx), &ob, &ib);
    if (ob.pos > 0) {
        env->SetByteArrayRegion(outputByteArray, outputStart, static_cast<jsize>(ob.pos), out.data());
        if (env->ExceptionCheck()) return kGenericError;
    }
    env->SetIntField(inputSlice, g_readStart, inputStart + static_cast<jint>(ib.pos));
    env->SetIntField(outputSlice, g_writeStart, outputStart + static_cast<jint>(ob.pos));
    return static_cast<jlong>(result);
}

JNIEXPORT jlong JNICALL Java_com_ensody_kompressor_zstd_ZstdWrapper_createCompressor(JNIEnv*, jobject) {
    return reinterpret_cast<jlong>(kzg_createCCtx());         // 0 if no GPU: Kotlin throws "Failed allocating zstd cctx"
}

JNIEXPORT void JNICALL Java_com_ensody_kompressor_zstd_ZstdWrapper_freeCompressor(JNIEnv*, jobject, jlong cctx) {
    kzg_freeCCtx(reinterpret_cast<kzg_CCtx*>(cctx));
}

JNIEXPORT jlong JNICALL Java_com_ensody_kompressor_zstd_ZstdWrapper_setParameter(JNIEnv*, jobject, jlong cctx, jint parameter, jint value) {
    return static_cast<jlong>(kzg_CCtx_setParameter(reinterpret_cast<kzg_CCtx*>(cctx), parameter, value));
}

JNIEXPORT jlong JNICALL Java_com_ensody_kompressor_zstd_ZstdWrapper_loadCompressorDictionary(JNIEnv* env, jobject, jlong cctx,
                                                                                             jbyteArray dictionary) {
    const jsize n = env->GetArrayLength(dictionary);
    std::vector<jbyte> buf(static_cast<size_t>(n > 0 ? n : 0));
    if (n > 0) {
        env->GetByteArrayRegion(dictionary, 0, n, buf.data());
        if (env->ExceptionCheck()) return kGenericError;
    }
    return static_cast<jlong>(kzg_CCtx_loadDictionary(reinterpret_cast<kzg_CCtx*>(cctx), buf.data(), buf.size()));
}

JNIEXPORT jlong JNICALL Java_com_ensody_kompressor_zstd_ZstdWrapper_compressStream(
        JNIEnv* env, jobject, jlong cctx, jobject inputSlice, jbyteArray inputByteArray, jint inputStart, jint inputEndExclusive,
        jobject outputSlice, jbyteArray outputByteArray, jint outputStart, jint outputEndExclusive, jboolean finish) {
    if (!resolve_fields(env)) return kGenericError;
    const jsize in_n = inputEndExclusive > inputStart ? inputEndExclusive - inputStart : 0;
    const jsize out_n = outputEndExclusive > outputStart ? outputEndExclusive - outputStart : 0;
    std::vector<jbyte> in(static_cast<size_t>(in_n)), out(static_cast<size_t>(out_n));
    if (in_n > 0) {
        env->GetByteArrayRegion(inputByteArray, inputStart, in_n, in.data());
        if (env->ExceptionCheck()) return kGenericError;
    }
    kzg_inBuffer ib = { in.data(), static_cast<size_t>(in_n), 0 };
    kzg_outBuffer ob = { out.data(), static_cast<size_t>(out_n), 0 };
    const size_t result = kzg_compressStream2(reinterpret_cast<kzg_CCtx*>(cctx), &ob, &ib, finish ? 2 : 0);
    if (ob.pos > 0) {
        env->SetByteArrayRegion(outputByteArray, outputStart, static_cast<jsize>(ob.pos), out.data());
        if (env->ExceptionCheck()) return kGenericError;
    }
    env->SetIntField(inputSlice, g_readStart, inputStart + static_cast<jint>(ib.pos));
    env->SetIntField(outputSlice, g_writeStart, outputStart + static_cast<jint>(ob.pos));
    return static_cast<jlong>(result);
}

JNIEXPORT jstring JNICALL Java_com_ensody_kompressor_zstd_ZstdWrapper_getErrorName(JNIEnv* env, jobject, jlong code) {
    const size_t c = static_cast<size_t>(code);
    if (!kzg_isError(c)) return nullptr;
    return env->NewStringUTF(kzg_getErrorName(c));
}

}  // extern "C"
