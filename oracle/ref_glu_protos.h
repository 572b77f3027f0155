// prototypes the reference's draw code expects from <GL/glu.h> (absent here); never called by the headless driver
#pragma once
#include <cstdint>
#ifdef __cplusplus
extern "C" {
#endif
typedef struct GLUquadric GLUquadric;
GLUquadric* gluNewQuadric(void);
void gluDeleteQuadric(GLUquadric*);
void gluQuadricDrawStyle(GLUquadric*, unsigned int);
void gluSphere(GLUquadric*, double, int, int);
void gluCylinder(GLUquadric*, double, double, double, int, int);
void gluDisk(GLUquadric*, double, double, int, int);
void gluPerspective(double, double, double, double);
void gluLookAt(double, double, double, double, double, double, double, double, double);
int gluProject(double, double, double, const double*, const double*, const int*, double*, double*, double*);
int gluUnProject(double, double, double, const double*, const double*, const int*, double*, double*, double*);
const unsigned char* gluErrorString(unsigned int);
#define GLU_FILL 100012
#define GLU_LINE 100011
#define GLU_SILHOUETTE 100013
#ifdef __cplusplus
}
#endif
