// TEST INFRASTRUCTURE (oracle/ref): empty bodies for the OpenGL-only virtuals whose real
// definitions live in ExternalLibrary/viewport.cpp:537-676 (GLUT is not available, and the
// viewport is out of scope: SURVEY.md section 2 row 14).
#include "std_first.h"
#include "ExternalLibrary/scene.h"
#include "ExternalLibrary/objects.h"
#include "ExternalLibrary/materials.h"
#include "ExternalLibrary/lights.h"
#include "ExternalLibrary/texture.h"

void Sphere::ViewportDisplay(const Material *) const {}
void Plane::ViewportDisplay(const Material *) const {}
void TriObj::ViewportDisplay(const Material *) const {}
void MtlBlinn::SetViewportMaterial(int) const {}
void GenLight::SetViewportParam(int, ColorA, ColorA, Point4) const {}
void PointLight::SetViewportLight(int) const {}
bool TextureFile::SetViewportTexture() const { return false; }
bool TextureChecker::SetViewportTexture() const { return false; }
