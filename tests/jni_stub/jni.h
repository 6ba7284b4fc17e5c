/* TEST INFRASTRUCTURE ONLY: the handful of JNI declarations kompressor_b200/csrc/kzg_jni.cpp uses, so
 * that the JNI shim can at least be syntax- and type-checked in an image without a JDK (tests/test_abi.py).
 * A real build uses the JDK's <jni.h> (kompressor_b200/build.py: build_jni). */
#ifndef KZG_TEST_JNI_STUB_H
#define KZG_TEST_JNI_STUB_H
#include <stdint.h>
typedef int32_t jint;
typedef int64_t jlong;
typedef int8_t jbyte;
typedef jint jsize;
typedef uint8_t jboolean;
class _jobject {};
typedef _jobject* jobject;
typedef jobject jclass;
typedef jobject jstring;
typedef jobject jarray;
typedef jarray jbyteArray;
struct _jfieldID;
typedef _jfieldID* jfieldID;
#define JNIEXPORT __attribute__((visibility("default")))
#define JNICALL
#define JNI_OK 0
#define JNI_ERR (-1)
#define JNI_VERSION_1_4 0x00010004
struct JNIEnv {
    jclass FindClass(const char*);
    jfieldID GetFieldID(jclass, const char*, const char*);
    jsize GetArrayLength(jarray);
    void GetByteArrayRegion(jbyteArray, jsize, jsize, jbyte*);
    void SetByteArrayRegion(jbyteArray, jsize, jsize, const jbyte*);
    jboolean ExceptionCheck();
    void SetIntField(jobject, jfieldID, jint);
    jstring NewStringUTF(const char*);
    jint GetVersion();
};
struct JavaVM {
    jint GetEnv(void**, jint);
};
#endif
